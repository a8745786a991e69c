"""GPU parity tests of the collimator stepping for pions and muons (mc_hms_coll / mc_shms_coll,
hms/mc_hms_coll.f, shms/mc_shms_coll.f, hms/pion_coll_absorb.f): with using_HMScoll / using_SHMScoll a pion that
enters the collimator material is not cut but stepped through it in 20 slices -- absorption, multiple
scattering, sampled energy loss, decay in flight -- and may punch through."""
import os
import re

import numpy as np
import pytest

from simc_gfortran_b200 import RunConfig, Simc, config_from_deck, load_optics_fixture
from tests.oracle_lib import transport_inputs
from tests.test_loop_gpu import RECON_LOOSE, accum_equal_exact
from tests.test_transport_gpu import compare

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("arm,decay", [(1, False), (1, True), (5, False), (5, True)])
def test_single_arm_with_collimator(oracle_with_optics, arm, decay):
    cfg = RunConfig()
    cfg.ctau = 780.4
    s = Simc(cfg, mode="strict")
    try:
        s.set_optics(load_optics_fixture(arm))
        n = 30000
        inp = transport_inputs(arm, n, seed=21, p_spec=2500.0, m2=139.57018 ** 2)
        inp[0] *= 0.6
        ref_out, ref_flags = oracle_with_optics.transport_batch(arm, inp, seed=17, decay=decay, coll=True, ctau=cfg.ctau)
        out, flags = s.transport_batch(arm, inp, 17, decay_flag=decay, using_coll=True)
        compare(out, flags, ref_out, ref_flags)
        # the same rows without the stepping: every accepted track is still accepted or lost in the collimator
        # material; some tracks the plain apertures cut now punch through
        plain, pflags = s.transport_batch(arm, inp, 17, decay_flag=decay, using_coll=False)
        coll_code = 19 if arm == 1 else 43
        slit = (1, 2, 3) if arm == 1 else (5, 6, 7)
        assert (flags == coll_code).sum() > 100                       # absorbed or ranged out
        assert not np.isin(flags, slit).any()                         # no slit cut with the stepping on
        through = np.isin(pflags, slit) & (flags == 0)
        assert through.sum() > 5                                      # punch-through pions reach the focal plane
    finally:
        s.close()


def test_electrons_are_not_stepped(oracle_with_optics):
    s = Simc(mode="strict")
    try:
        s.set_optics(load_optics_fixture(1))
        inp = transport_inputs(1, 5000, seed=3)
        a, fa = s.transport_batch(1, inp, 5, using_coll=True)
        b, fb = s.transport_batch(1, inp, 5, using_coll=False)
        assert np.array_equal(fa, fb) and np.array_equal(a, b)       # mc_hms.f:206: pion / muon masses only
    finally:
        s.close()


def test_loop_with_shms_collimator(oracle_with_optics):
    """C3 (H(e,e'pi+)n, pions in the SHMS) with using_SHMScoll = 1: counters incl. the slit STOP counters, which
    the stepping bumps once per slice spent in the material (mc_shms_coll.f)."""
    txt = open(os.path.join(ROOT, "decks", "c3_eepi_hydrogen_hms_shms.inp")).read()
    txt, n = re.subn(r"(begin parm simulate\n)", r"\1  using_SHMScoll = 1\n", txt, count=1)
    assert n == 1
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "deck.inp")
        open(path, "w").write(txt)
        cfg = config_from_deck(path)[0]
    assert cfg.using_SHMScoll and not cfg.using_HMScoll
    sim = Simc(cfg, mode="strict")
    try:
        for arm in (1, 5):
            sim.set_optics(load_optics_fixture(arm))
        n = 60000
        ref = oracle_with_optics.run(cfg, 0, n, 4, threads=8)
        acc = sim.accum_clear()
        sim.run(0, n, 4, acc)
        accum_equal_exact(acc, ref)
        stop = np.ctypeslib.as_array(acc.stop)[1]
        assert stop[2 + 43] > 50                                      # shmsSTOP_coll
        assert stop[2 + 5] + stop[2 + 6] + stop[2 + 7] > stop[2 + 43]  # slices in material >> events lost
        a, b = acc.wtcontribute.value(), ref.wtcontribute.value()
        assert abs(a - b) <= RECON_LOOSE * abs(b)
    finally:
        sim.close()
