import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from simc_gfortran_b200 import Simc, config_from_deck, load_optics_fixture
from tests.oracle_lib import Oracle
cfg = config_from_deck('decks/c1_eep_hydrogen_hms_shms.inp')[0]
orc = Oracle()
for arm in (1, 5): orc.set_optics(load_optics_fixture(arm))
n = 50000
rec, stage = orc.event_batch(cfg, 0, n, 5)
m = stage >= 1
Ein, eE = rec[10][m], rec[11][m]
rng = np.random.default_rng(3)
k = m.sum()
Mp = 938.27231
cth = 1.0 - (Ein / eE - 1.0) * Mp / Ein
eth = np.arccos(cth)
phi = rng.uniform(4.6, 4.8, k)
ue = np.stack([np.sin(eth) * np.cos(phi), np.sin(eth) * np.sin(phi), np.cos(eth)])
nu = Ein - eE
q = np.sqrt(2 * Ein * eE * (1 - ue[2]) + nu * nu)
up = np.stack([-eE * ue[0] / q, -eE * ue[1] / q, (Ein - eE * ue[2]) / q])
pE = np.sqrt(q * q + Mp * Mp)
emax = rng.uniform(20.0, 1200.0, k)
emin = np.where(rng.uniform(size=k) < 0.5, rng.uniform(-50.0, 0.0, k), rng.uniform(0.0, 0.9, k) * emax)
inp = np.stack([Ein, eE, eth, ue[0], ue[1], ue[2], pE, q, up[0], up[1], up[2], rng.uniform(2e-3, 3e-2, k),
                rng.uniform(5e-3, 5e-2, k), rng.uniform(0.0, 1.0, k) * emax, emin, emax])
ref = orc.radc_batch(cfg, inp)
sim = Simc(cfg, mode='strict')
out = sim.radc_batch(inp)
err = np.abs(out - ref) / np.maximum(np.abs(ref), 1e-300)
print('per-row max', err.max(axis=1))
print('per-row frac>1e-12', (err > 1e-12).mean(axis=1))
j = int(np.argmax(err[7]))
print('worst col', j, 'inp', inp[:, j], 'ref', ref[:, j], 'out', out[:, j])
