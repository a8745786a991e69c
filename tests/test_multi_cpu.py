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


def test_pack_unpack_round_trip(built_lib):
    from simc_gfortran_b200.lib import Accum
    from simc_gfortran_b200.multi import pack, unpack
    a = Accum()
    a.ntried = 123; a.wtcontribute.lo = 2 ** 64 - 5; a.wtcontribute.hi = -3; a.sumerr[2].lo = 7; a.sumerr[2].hi = -1
    a.hist_n[1][2][3] = 9; a.contrib[4].lo = -1.5; a.contrib[4].hi = 2.5e9; a.stop[1][7] = 11
    b = unpack(a, *pack(a))
    assert bytes(a) == bytes(b)
    s, mn, mx = pack(a)
    c = unpack(a, 2 * s, mn, mx)          # what a 2-rank sum of identical accumulators gives
    assert c.ntried == 246 and c.stop[1][7] == 22
    v = ((int(a.wtcontribute.hi) << 64) + int(a.wtcontribute.lo)) * 2
    assert ((int(c.wtcontribute.hi) << 64) + int(c.wtcontribute.lo)) == v
