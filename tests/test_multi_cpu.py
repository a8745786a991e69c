"""The N>1 path on CPU: two gloo ranks reduce their accumulators exactly (sharding by try range +
one all-reduce of integers).  The per-rank accumulators come from the oracle here (no
GPU in this container); the reduction code is the product's (simc_gfortran_b200.multi)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import torch.distributed as dist
from simc_gfortran_b200 import config_from_deck, load_optics_fixture
from simc_gfortran_b200.multi import allreduce_accum
from tests.oracle_lib import Oracle
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
cfg = config_from_deck(os.path.join(%(root)r, "decks", "c1_eep_hydrogen_hms_shms.inp"))[0]
orc = Oracle()
for arm in (1, 5):
    orc.set_optics(load_optics_fixture(arm))
n = 6000
mine = orc.run(cfg, rank * n, n, 31, threads=2)
total = allreduce_accum(mine)
if rank == 0:
    whole = orc.run(cfg, 0, world * n, 31, threads=2)
    assert bytes(total) == bytes(whole), "sharded + all-reduced accumulators differ from the single-rank run"
    assert total.ntried == world * n and total.nsuccess > 1000
    print("OK", total.ntried, total.nsuccess)
dist.destroy_process_group()
'''


def test_two_rank_allreduce_is_exact(tmp_path, built_lib):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "OK 12000" in out.stdout


def test_accum_merge_is_exact(built_lib):
    """simc_b200_accum_merge: counters add, 128-bit sums carry, ranges widen; different quanta are refused."""
    import pytest
    from simc_gfortran_b200.lib import Accum, SimcError, accum_merge
    a = Accum()
    a.ntried = 123; a.nonfinite = 2; a.unsupported = 1
    a.wtcontribute.lo = 2 ** 64 - 5; a.wtcontribute.hi = -3; a.wtcontribute.qexp = -70
    a.sumerr[2].lo = 7; a.sumerr[2].hi = -1
    a.hist_n[1][2][3] = 9; a.contrib[4].lo = -1.5; a.contrib[4].hi = 2.5e9; a.stop[1][7] = 11; a.transp_calls[1][47] = 5
    b = Accum.from_buffer_copy(bytes(a))
    b.contrib[4].lo = -2.5; b.contrib[4].hi = 1.0
    c = accum_merge(Accum.from_buffer_copy(bytes(a)), b)
    assert c.ntried == 246 and c.stop[1][7] == 22 and c.nonfinite == 4 and c.unsupported == 2 and c.transp_calls[1][47] == 10
    assert c.hist_n[1][2][3] == 18 and c.contrib[4].lo == -2.5 and c.contrib[4].hi == 2.5e9
    v = ((int(a.wtcontribute.hi) << 64) + int(a.wtcontribute.lo)) * 2
    assert ((int(c.wtcontribute.hi) << 64) + int(c.wtcontribute.lo)) == v and c.wtcontribute.qexp == -70
    assert ((int(c.sumerr[2].hi) << 64) + int(c.sumerr[2].lo)) == 2 * (7 - 2 ** 64)
    # an empty sum takes the other's quantum; two non-empty sums on different quanta cannot be added
    e = accum_merge(Accum(), a)
    assert bytes(e.wtcontribute) == bytes(a.wtcontribute)
    b.wtcontribute.qexp = -60
    with pytest.raises(SimcError):
        accum_merge(Accum.from_buffer_copy(bytes(a)), b)


NGEN_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import torch.distributed as dist
from simc_gfortran_b200 import config_from_deck, load_optics_fixture, Accum
from simc_gfortran_b200.multi import run_until_successes
from tests.oracle_lib import Oracle
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
cfg = config_from_deck(os.path.join(%(root)r, "decks", "c1_eep_hydrogen_hms_shms.inp"))[0]
orc = Oracle()
for arm in (1, 5):
    orc.set_optics(load_optics_fixture(arm))
run = lambda first, n: orc.run(cfg, first, n, 77, threads=2)
new = lambda: orc.run(cfg, 0, 0, 77, threads=1)
ngen = 700
total, cut = run_until_successes(run, new, ngen, 1500)
assert total.nsuccess == ngen and total.ntried == cut, (total.nsuccess, total.ntried, cut)
if rank == 0:
    whole = run(0, cut)                       # the single-process loop stopped at the same try
    assert bytes(total) == bytes(whole), "multi-rank ngen > 0 run differs from the single-rank one"
    assert run(0, cut - 1).nsuccess == ngen - 1, "the last try must be the ngen-th success (simc.f:346-350)"
    print("OK", cut, total.nsuccess)
dist.destroy_process_group()
'''


def test_two_rank_ngen_positive_stops_at_the_same_try(tmp_path, built_lib):
    """ngen > 0 over two gloo ranks (simc.f:346-350): the cut try index and every accumulator equal the
    single-rank run's."""
    script = tmp_path / "worker_ngen.py"
    script.write_text(NGEN_WORKER % {"root": ROOT})
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29535", str(script)],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "OK" in out.stdout
