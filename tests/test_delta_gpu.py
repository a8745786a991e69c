"""GPU parity tests of H(e,e'p)pi0 ("doing_delta", dbase.f:182-190, event.f:680-684, physics_delta.f): the proton
energy is one of the two roots of the two-body quadratic, chosen by a coin toss, and the weight is phase space
times the virtual-photon flux.  Kinematics of the reference's infiles/test_delta_h.inp (SOS electron, HMS proton),
radiation off: generate_rad sets no photon-energy limits for this reaction (radc.f:249-294)."""
import os

import numpy as np
import pytest

from simc_gfortran_b200 import Simc, SimcError, config_from_deck, load_optics_fixture
from tests.test_loop_gpu import LOOSE, RECON_LOOSE, SCALE, accum_equal_exact, rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECK = os.path.join(ROOT, "decks", "e1_eep_pi0_hydrogen_sos_hms.inp")
SC = SCALE.copy()
SC[50] = 1e-12
SC[51] = 1e-3
SC[53] = 1.0
SC[55] = 1e3


@pytest.fixture(scope="module")
def case(oracle_with_optics):
    cfg = config_from_deck(DECK)[0]
    s = Simc(cfg, mode="strict")
    for arm in (1, 2):
        s.set_optics(load_optics_fixture(arm))
    yield cfg, s, oracle_with_optics
    s.close()


def test_config(case):
    cfg = case[0]
    assert cfg.doing_delta == 1 and cfg.doing_pion == 0 and cfg.doing_eep == 0
    assert abs(cfg.Mh - 938.27231) < 1e-9 and abs(cfg.targ.Mrec_struck - 139.57018) < 1e-9     # dbase.f:183,297-299
    assert cfg.cuts_Em.min == -1.0e6 and cfg.cuts_Em.max == 1.0e6                                # dbase.f:534-538


def test_event_records(case):
    cfg, sim, orc = case
    n = 60000
    ref, ref_stage = orc.event_batch(cfg, 0, n, 7)
    rec, stage = sim.event_batch(0, n, 7)
    assert np.array_equal(stage, ref_stage)
    for k in (0, 2, 3, 4):
        assert np.array_equal(rec[k], ref[k])            # try index, draws consumed (the coin toss included), stop codes
    names = sim.event_field_names()
    done = stage == 4
    assert done.sum() > 500
    for k in (1, 5, 6, 9, 44, 45, 46, 52, 53):
        e = rel_err(rec[k][done], ref[k][done], SC[k])
        assert e.max() <= RECON_LOOSE, (names[k], float(e.max()))
    for k in (48, 49, 50, 51, 54, 55):
        e = rel_err(rec[k][done], ref[k][done], SC[k])
        assert e.max() <= LOOSE, (names[k], float(e.max()))


def test_accumulators(case):
    cfg, sim, orc = case
    n = 60000
    ref = orc.run(cfg, 0, n, 4, threads=8)
    acc = sim.accum_clear()
    sim.run(0, n, 4, acc)
    accum_equal_exact(acc, ref)
    assert acc.unsupported == ref.unsupported == 0 and acc.nsuccess > 500
    a, b = acc.wtcontribute.value(), ref.wtcontribute.value()
    assert abs(a - b) <= RECON_LOOSE * abs(b)


def test_ntuple_rows(case):
    cfg, sim, orc = case
    n = 30000
    ref, ref_tries = orc.ntuple_batch(cfg, 0, n, 5)
    rows, tries = sim.ntuple_batch(0, n, 5)
    assert rows.shape[1] == ref.shape[1] == 53               # NtupleInit.f:101: the pion/kaon/delta layout
    assert np.array_equal(tries, ref_tries) and len(rows) > 300
    scale = np.maximum(np.abs(ref).max(axis=0), 1e-30)
    err = np.abs(rows - ref) / np.maximum(np.abs(ref), 1e-3 * scale[None, :])
    assert err.max() <= RECON_LOOSE, (int(np.argmax(err.max(axis=0))), float(err.max()))
    # both roots of the quadratic are thrown; the HMS window keeps the high-momentum one
    assert np.all(np.abs(rows[:, 20]) < 60)


def test_refusals(case):
    cfg = case[0]
    rad = type(cfg).from_buffer_copy(bytes(cfg))
    rad.using_rad = 1
    s = Simc(rad, mode="strict")
    try:
        for arm in (1, 2):
            s.set_optics(load_optics_fixture(arm))
        with pytest.raises(SimcError) as e:
            s.run(0, 1000, 1, s.accum_clear())
        assert "radc.f" in str(e.value)
    finally:
        s.close()
