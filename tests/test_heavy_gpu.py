"""GPU parity tests of the event loop for C2 = C-12 A(e,e'p) with the Benhar spectral function
(HMS electron + SOS proton): per-try records and exact accumulators of libsimc_b200 against the
CPU oracle on the same counter-based random stream, plus the spectral-function lookup on dumped
(Em, Pm) vectors.  Tolerances as in tests/test_loop_gpu.py."""
import os

import numpy as np
import pytest

from simc_gfortran_b200 import Simc, SimcError, config_from_deck, load_optics_fixture
from tests.test_loop_gpu import LOOSE, RECON_LOOSE, RTOL, SCALE, accum_equal_exact, rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECK = os.path.join(ROOT, "decks", "c2_eep_carbon_hms_sos.inp")
SC = SCALE.copy()
SC[5] = 1e-12        # weights ~1e-9
SC[6] = SC[9] = 1.0  # deForest cross sections ~1e2


@pytest.fixture(scope="module")
def sf():
    return np.load(os.path.join(ROOT, "simc_gfortran_b200", "data", "benharsf_12.npz"))


@pytest.fixture(scope="module")
def case(oracle_with_optics, sf):
    cfg = config_from_deck(DECK)[0]
    assert cfg.doing_heavy and cfg.use_benhar_sf and cfg.w_ref < 1e-6
    s = Simc(cfg, mode="strict")
    for arm in (1, 2):
        s.set_optics(load_optics_fixture(arm))
    s.set_sf_table(sf["pm"], sf["em"], sf["sf_proton"])
    oracle_with_optics.set_sf_table(sf["pm"], sf["em"], sf["sf_proton"])
    yield cfg, s, oracle_with_optics
    s.close()


def test_event_records(case):
    cfg, sim, orc = case
    n = 40000
    ref, ref_stage = orc.event_batch(cfg, 0, n, 3)
    rec, stage = sim.event_batch(0, n, 3)
    assert np.array_equal(stage, ref_stage), f"{(stage != ref_stage).sum()} tries end at a different stage"
    for k in (0, 2, 3, 4):
        assert np.array_equal(rec[k], ref[k]), sim.event_field_names()[k]
    names = sim.event_field_names()
    gen_ok = stage >= 1
    for k in (8, 13, 14, 17, 18, 26, 27, 28, 29):
        e = rel_err(rec[k][gen_ok], ref[k][gen_ok], SC[k])
        assert e.max() <= RTOL, (names[k], float(e.max()))
    groups = (
        ([7] + list(range(10, 32)) + [35, 36, 37, 47], stage >= 1, LOOSE),
        ([32, 33, 34], stage >= 2, LOOSE),
        ([41, 42, 43], stage >= 2, RECON_LOOSE),
        ([38, 39, 40], stage >= 3, RECON_LOOSE),
        ([1, 5, 6, 9, 44, 45, 46], stage == 4, RECON_LOOSE),
    )
    for fields, mask, tol in groups:
        for k in fields:
            e = rel_err(rec[k][mask], ref[k][mask], SC[k])
            assert e.max() <= tol, (names[k], float(e.max()))
    assert (stage == 4).sum() > 1000 and (stage == 0).sum() > 1000 and (stage == 3).sum() > 50   # 3: no strength at (Em,Pm)


def test_accumulators_against_oracle(case):
    cfg, sim, orc = case
    n = 60000
    ref = orc.run(cfg, 0, n, 4, threads=8)
    acc = sim.accum_clear()
    sim.run(0, n, 4, acc)
    accum_equal_exact(acc, ref)
    assert acc.unsupported == ref.unsupported == 0 and acc.nsuccess > 1500
    for f in ("wtcontribute", "sum_sigcc"):
        a, b = getattr(acc, f).value(), getattr(ref, f).value()
        assert abs(a - b) <= RECON_LOOSE * abs(b), f
        assert getattr(acc, f).qexp == getattr(ref, f).qexp


def test_needs_the_table():
    cfg = config_from_deck(DECK)[0]
    s = Simc(cfg)
    try:
        for arm in (1, 2):
            s.set_optics(load_optics_fixture(arm))
        with pytest.raises(SimcError) as e:
            s.run(0, 10, 1, s.accum_clear())
        assert "spectral function" in str(e.value)
    finally:
        s.close()
