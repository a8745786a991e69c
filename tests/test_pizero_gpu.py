"""GPU parity tests of H(e,e'pi0)p -> gamma gamma into a calorimeter arm ("doing_pizero", hadron_arm = 8: pizero_decay.f,
calo/mc_calo.f, simc.f:1489-1564, 1605-1609; dbase.f:140, 330-347; event.f:899-901, 1464-1490).  The pi0 is generated
like any exclusive pion; complete_ev throws the decay angles (two random numbers per pass), montecarlo drifts both
photons to the calorimeter front and asks for pizero_ngamma of them inside its face.  Kinematics of the reference's
infiles/nps_excl_pi0_test.inp (HMS electron, NPS on the SHMS side), radiative tails on as in that deck."""
import os

import numpy as np
import pytest

from simc_gfortran_b200 import Simc, SimcError, config_from_deck, load_optics_fixture
from tests.test_loop_gpu import LOOSE, RECON_LOOSE, SCALE, accum_equal_exact, rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECK = os.path.join(ROOT, "decks", "z1_eepi0_hydrogen_hms_nps.inp")
SC = SCALE.copy()
for k in (5, 6):
    SC[k] = 1e-9
SC[50] = 1e-3
SC[51] = 1e3
SC[53] = 1.0
SC[55] = 1e3


def variant(cfg, **kw):
    c = type(cfg).from_buffer_copy(bytes(cfg))
    for k, v in kw.items():
        setattr(c, k, v)
    return c


@pytest.fixture(scope="module", params=[2, 1])
def case(request, oracle_with_optics):
    cfg = variant(config_from_deck(DECK)[0], pizero_ngamma=request.param)
    s = Simc(cfg, mode="strict")
    s.set_optics(load_optics_fixture(1))          # the electron arm; a calorimeter has no maps
    yield cfg, s, oracle_with_optics
    s.close()


def test_config(case):
    cfg = case[0]
    assert cfg.doing_pizero == 1 and cfg.doing_pion == 1 and cfg.doing_hydpi == 1 and cfg.hadron_arm == 8
    assert abs(cfg.Mh - 134.9766) < 1e-12 and abs(cfg.targ.Mrec_struck - 938.27231) < 1e-9     # dbase.f:140, 331-333
    assert cfg.drift_to_cal == 300.0 and abs(cfg.spec_p.phi - np.pi / 2) < 1e-15                # dbase.f:259-261


def test_event_records(case):
    cfg, sim, orc = case
    n = 60000
    ref, ref_stage = orc.event_batch(cfg, 0, n, 7)
    rec, stage = sim.event_batch(0, n, 7)
    assert np.array_equal(stage, ref_stage), f"{(stage != ref_stage).sum()} tries end at a different stage"
    for k in (0, 2, 3, 4):
        assert np.array_equal(rec[k], ref[k]), sim.event_field_names()[k]      # try, draws, stop codes of both arms
    names = sim.event_field_names()
    gen, done = stage >= 1, stage == 4
    assert gen.sum() > 10000 and done.sum() > 100 and (rec[3] == 1).sum() > 100 and (rec[3] == 2).sum() > 100
    for k in (7, 8, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 26, 27, 28, 29, 30, 31, 35, 36, 37):
        e = rel_err(rec[k][gen], ref[k][gen], SC[k])
        assert e.max() <= LOOSE, (names[k], float(e.max()))
    # no pi0 reconstruction: the recon quantities of the hadron arm are its SP quantities (simc.f:1605-1609)
    for a, b in ((41, 35), (42, 36), (43, 37)):
        assert np.array_equal(rec[a][done], rec[b][done])
    for k in (1, 5, 6, 9, 38, 39, 40, 44, 45, 46, 48, 49, 50, 51, 53, 55):
        e = rel_err(rec[k][done], ref[k][done], SC[k])
        assert e.max() <= RECON_LOOSE, (names[k], float(e.max()))


def test_accumulators(case):
    cfg, sim, orc = case
    n = 60000
    ref = orc.run(cfg, 0, n, 4, threads=8)
    acc = sim.accum_clear()
    sim.run(0, n, 4, acc)
    accum_equal_exact(acc, ref)
    assert acc.unsupported == ref.unsupported and acc.nsuccess > 100
    a, b = acc.wtcontribute.value(), ref.wtcontribute.value()
    assert abs(a - b) <= RECON_LOOSE * abs(b)


def test_ntuple_rows(case):
    cfg, sim, orc = case
    n = 60000
    ref, ref_tries = orc.ntuple_batch(cfg, 0, n, 5)
    rows, tries = sim.ntuple_batch(0, n, 5)
    assert rows.shape[1] == ref.shape[1] == 65               # NtupleInit.f:101-184: the pion layout + 12 photon columns
    assert np.array_equal(tries, ref_tries) and len(rows) > 100
    scale = np.maximum(np.abs(ref).max(axis=0), 1e-30)
    scale[[53, 54, 59, 60]] = 100.0                          # hit positions (cm), -1e10 when a photon missed
    err = np.abs(rows - ref) / np.maximum(np.abs(ref), 1e-3 * scale[None, :])
    assert err.max() <= RECON_LOOSE, (int(np.argmax(err.max(axis=0))), float(err.max()))
    # the two photons carry the pi0: energies add up to the vertex pi0 energy, invariant mass = Mpi0
    g1, g2 = rows[:, 55:59], rows[:, 61:65]
    tot = g1 + g2
    m2 = tot[:, 0] ** 2 - (tot[:, 1:] ** 2).sum(axis=1)
    assert np.allclose(np.sqrt(m2), 134.9766, rtol=1e-9)
    assert np.allclose(np.sqrt((g1[:, 1:] ** 2).sum(axis=1)), g1[:, 0], rtol=1e-12)          # massless
    inside = (np.abs(rows[:, [53, 59]]) <= 36.9).all(axis=1) & (np.abs(rows[:, [54, 60]]) <= 30.75).all(axis=1)
    if cfg.pizero_ngamma == 2:
        assert inside.all()
    else:
        missed = (rows[:, 53] == -1.0e10) | (rows[:, 59] == -1.0e10)
        assert missed.any() and (inside | missed).all()


def test_refusals(case):
    cfg = case[0]
    bad = variant(cfg, electron_arm=7)
    s = Simc(bad, mode="strict")
    try:
        with pytest.raises(SimcError) as e:
            s.run(0, 1000, 1, s.accum_clear())
        assert "calorimeter" in str(e.value)
    finally:
        s.close()
