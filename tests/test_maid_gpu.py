"""GPU parity tests of peepi's MAID-2007 branch (sigmaid, physics_pion.f:131-154, 577-728): H(e,e'pi+)n in the
Delta region (W = 1.16 GeV, the reference's test_pion_h.inp setting, HMS + SOS).  With the table every
contributing event takes the blended cross section; without it they are counted in `unsupported`."""
import os

import numpy as np
import pytest

from simc_gfortran_b200 import Simc, config_from_deck, load_optics_fixture
from tests.oracle_lib import load_maid_fixture, write_maid_file
from tests.test_loop_gpu import LOOSE, RECON_LOOSE, SCALE, accum_equal_exact, rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECK = os.path.join(ROOT, "decks", "p1_eepi_hydrogen_lowW_hms_sos.inp")
SC = SCALE.copy()
SC[50] = 1e-12
SC[51] = 1e-3
SC[53] = 1.0
SC[55] = 1e3


@pytest.fixture(scope="module")
def case(oracle_with_optics):
    cfg = config_from_deck(DECK)[0]
    tbl = load_maid_fixture(3)
    oracle_with_optics.set_maid_table(3, tbl)
    s = Simc(cfg, mode="strict")
    for arm in (1, 2):
        s.set_optics(load_optics_fixture(arm))
    s.set_maid_table(3, tbl)
    yield cfg, s, oracle_with_optics
    s.close()
    oracle_with_optics.set_maid_table(3, None)


def test_event_records(case):
    cfg, sim, orc = case
    n = 60000
    ref, ref_stage = orc.event_batch(cfg, 0, n, 7)
    rec, stage = sim.event_batch(0, n, 7)
    assert np.array_equal(stage, ref_stage)
    for k in (0, 2, 3, 4):
        assert np.array_equal(rec[k], ref[k])
    names = sim.event_field_names()
    done = stage == 4
    assert done.sum() > 200
    assert np.all(rec[54][done] < 1500.0)                      # W: pure MAID below 1.5 GeV (fac1 = 0)
    for k in (1, 5, 6, 9, 44, 45, 46, 52, 53):
        e = rel_err(rec[k][done], ref[k][done], SC[k])
        assert e.max() <= RECON_LOOSE, (names[k], float(e.max()))
    for k in (48, 49, 50, 51, 54, 55):
        e = rel_err(rec[k][done], ref[k][done], SC[k])
        assert e.max() <= LOOSE, (names[k], float(e.max()))


def test_accumulators_and_unsupported_counter(case):
    cfg, sim, orc = case
    n = 60000
    ref = orc.run(cfg, 0, n, 4, threads=8)
    acc = sim.accum_clear()
    sim.run(0, n, 4, acc)
    accum_equal_exact(acc, ref)
    assert acc.unsupported == ref.unsupported == 0 and acc.nsuccess > 200
    a, b = acc.wtcontribute.value(), ref.wtcontribute.value()
    assert abs(a - b) <= RECON_LOOSE * abs(b)
    # without the table: same events, every one of them flagged, and a different (parametrisation-only) weight
    s2 = Simc(cfg, mode="strict")
    try:
        for arm in (1, 2):
            s2.set_optics(load_optics_fixture(arm))
        acc2 = s2.accum_clear()
        s2.run(0, n, 4, acc2)
        assert acc2.nsuccess == acc.nsuccess and acc2.unsupported == acc2.ncontribute > 0
        assert abs(acc2.wtcontribute.value() / a - 1.0) > 0.05
    finally:
        s2.close()


def test_file_reader(case, tmp_path):
    cfg, sim, orc = case
    path = str(tmp_path / "maidpipn.dat")
    write_maid_file(load_maid_fixture(3), path)
    a = sim.accum_clear()
    sim.run(0, 20000, 2, a)
    s2 = Simc(cfg, mode="strict")
    try:
        for arm in (1, 2):
            s2.set_optics(load_optics_fixture(arm))
        s2.load_maid_file(3, path)
        b = s2.accum_clear()
        s2.run(0, 20000, 2, b)
        assert bytes(a) == bytes(b)
    finally:
        s2.close()
