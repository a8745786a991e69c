"""GPU parity of the target-field tracking (csrc/field.cuh against oracle/field.cpp; trg_track.f): track_from_tgt on
identical dumped vectors through simc_b200_field_batch, for the measured map (set from arrays and read from a file by
the library's own reader), the uniform test field and no field; 1e-12 on the image track, `ok` flags identical."""
import os

import numpy as np
import pytest

from simc_gfortran_b200 import Simc, SimcError, config_from_deck
from tests.oracle_lib import load_field_fixture, write_field_file

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECK = os.path.join(ROOT, "decks", "c1_eep_hydrogen_hms_shms.inp")


def vectors(n, seed):
    rng = np.random.default_rng(seed)
    return np.array([rng.uniform(-1, 1, n), rng.uniform(-1.5, 1.5, n), rng.uniform(-3, 3, n), rng.uniform(-0.08, 0.08, n),
                     rng.uniform(-0.05, 0.05, n), rng.choice([-1.0, 1.0], n) * rng.uniform(300, 6000, n),
                     rng.choice([0.51099906, 139.57018, 493.677, 938.27231], n)])


@pytest.fixture(scope="module")
def sim():
    s = Simc(config_from_deck(DECK)[0], mode="strict")
    yield s
    s.close()


@pytest.mark.parametrize("theta", [80.0, -100.0, 12.5, 0.0])
def test_measured_map(sim, oracle, theta, tmp_path):
    bz, br = load_field_fixture()
    inp = vectors(20000, 3)
    for how in ("arrays", "file"):
        if how == "arrays":
            sim.set_field_map(bz, br)
        else:
            path = str(tmp_path / "trg_field_map.dat")
            write_field_file(bz, br, path)
            sim.load_field_file(path)
        for spect in (-1, 1):
            oracle.set_field_map(bz, br, theta, theta)
            ref = oracle.field_batch(spect, inp)
            out = sim.field_batch(spect, theta, inp)
            assert np.array_equal(out[5], ref[5]) and ref[5].mean() > 0.99
            ok = ref[5] == 1
            err = np.abs(out[:5, ok] - ref[:5, ok]) / np.maximum(np.abs(ref[:5, ok]), [[1.0], [1.0], [1.0], [1e-2], [1e-2]])
            assert err.max() <= 1e-12, float(err.max())
    # low momenta curl up in the field: tracks that never reach z = 100 cm come back with ok = 0 from both
    slow = vectors(4000, 9)
    slow[5] = np.sign(slow[5]) * np.abs(slow[5]) / 40.0
    oracle.set_field_map(bz, br, theta, theta)
    ref = oracle.field_batch(1, slow)
    out = sim.field_batch(1, theta, slow)
    assert np.array_equal(out[5], ref[5])
    if abs(theta) > 45:
        assert (ref[5] == 0).any()


def test_uniform_and_zero_field(sim, oracle, tmp_path):
    inp = vectors(5000, 4)
    sim.set_field_map(None, None)
    oracle.set_field_map(None, None, 30.0, 30.0)
    ref, out = oracle.field_batch(1, inp), sim.field_batch(1, 30.0, inp)
    assert np.array_equal(out[5], ref[5])
    ok = ref[5] == 1
    assert np.abs(out[:5, ok] - ref[:5, ok]).max() <= 1e-11
    sim.load_field_file("0")                              # trgInit's "no field" option
    out = sim.field_batch(-1, 30.0, inp)
    assert np.all(out[5] == 1) and np.allclose(out[2], 100.0, atol=1e-9)
    assert np.allclose(out[0], inp[0] + inp[3] * (100.0 - inp[2]), atol=1e-9) and np.allclose(out[3], inp[3], atol=1e-13)
    sim.load_field_file(" ")                              # blank name: the uniform test field again
    out2 = sim.field_batch(1, 30.0, inp)
    assert np.array_equal(out2, sim.field_batch(1, 30.0, inp)) and np.abs(out2[:5, ok] - ref[:5, ok]).max() <= 1e-11


def test_needs_a_map():
    s = Simc(config_from_deck(DECK)[0], mode="strict")
    try:
        with pytest.raises(SimcError) as e:
            s.field_batch(1, 0.0, vectors(4, 1))
        assert "field map" in str(e.value)
    finally:
        s.close()


# ---- the whole loop with using_tgt_field: both arms tracked through the target's field -------------------------------
POLDECK = os.path.join(ROOT, "decks", "w1_poltar_eepi_hydrogen_hms_shms.inp")


@pytest.fixture(scope="module")
def polcase(oracle_with_optics):
    from simc_gfortran_b200 import load_optics_fixture
    cfg = config_from_deck(POLDECK)[0]
    bz, br = load_field_fixture()
    oracle_with_optics.set_field_map(bz, br, 0.0, 0.0)        # the run takes its angles from the deck (simc.f:120-156)
    s = Simc(cfg, mode="strict")
    for arm in (1, 5):
        s.set_optics(load_optics_fixture(arm))
    s.set_field_map(bz, br)
    yield cfg, s, oracle_with_optics
    s.close()


def test_loop_with_field_records(polcase):
    from tests.test_loop_gpu import LOOSE, RECON_LOOSE, SCALE, rel_err
    cfg, sim, orc = polcase
    assert cfg.using_tgt_field == 1 and abs(cfg.targ_Bangle - np.radians(80.0)) < 1e-15 and cfg.sign_hadron == 1.0
    n = 40000
    ref, ref_stage = orc.event_batch(cfg, 0, n, 7)
    rec, stage = sim.event_batch(0, n, 7)
    assert np.array_equal(stage, ref_stage), f"{(stage != ref_stage).sum()} tries end at a different stage"
    for k in (0, 2, 3, 4):
        assert np.array_equal(rec[k], ref[k]), sim.event_field_names()[k]
    names = sim.event_field_names()
    done = stage == 4
    assert done.sum() > 100
    sc = SCALE.copy()
    for k in (5, 6):
        sc[k] = 1e-9
    # SP quantities: the generated ones (the field does not touch them); recon: iterated against the field
    for fields, mask, tol in (((32, 33, 34), stage >= 2, LOOSE), ((35, 36, 37), stage >= 1, LOOSE),
                              ((41, 42, 43), stage >= 2, RECON_LOOSE), ((38, 39, 40, 1, 5, 6, 44, 45, 46), done, RECON_LOOSE)):
        for k in fields:
            e = rel_err(rec[k][mask], ref[k][mask], sc[k])
            assert e.max() <= tol, (names[k], float(e.max()))
    # the field matters: the same tries without it end differently and reconstruct differently
    off = type(cfg).from_buffer_copy(bytes(cfg))
    off.using_tgt_field = 0
    from simc_gfortran_b200 import load_optics_fixture
    s0 = Simc(off, mode="strict")
    try:
        for arm in (1, 5):
            s0.set_optics(load_optics_fixture(arm))
        rec0, stage0 = s0.event_batch(0, n, 7)
    finally:
        s0.close()
    assert (stage0 != stage).mean() > 0.01
    # reconstruction against the field recovers the vertex angles: resolution of the electron's yptar stays at the
    # mrad level (without track_to_tgt it would be off by the field's bend, tens of mrad)
    assert np.std(rec[39][done] - rec[33][done]) < 3e-3


def test_loop_with_field_accumulators(polcase):
    from tests.test_loop_gpu import RECON_LOOSE, accum_equal_exact
    cfg, sim, orc = polcase
    n = 40000
    ref = orc.run(cfg, 0, n, 4, threads=8)
    acc = sim.accum_clear()
    sim.run(0, n, 4, acc)
    accum_equal_exact(acc, ref)
    assert acc.nsuccess > 100
    a, b = acc.wtcontribute.value(), ref.wtcontribute.value()
    assert abs(a - b) <= RECON_LOOSE * abs(b)


def test_loop_with_field_ntuple(polcase):
    """The pion layout with the eight polarised-target columns (results_write.f:154-162): theta_tarq, phi_targ, beta,
    phi_s, phi_c of the reconstructed event, beta, phi_s, phi_c of the vertex."""
    from tests.test_loop_gpu import RECON_LOOSE
    cfg, sim, orc = polcase
    n = 40000
    ref, ref_tries = orc.ntuple_batch(cfg, 0, n, 5)
    rows, tries = sim.ntuple_batch(0, n, 5)
    assert rows.shape[1] == ref.shape[1] == 61 and np.array_equal(tries, ref_tries) and len(rows) > 100
    scale = np.maximum(np.abs(ref).max(axis=0), 1e-30)
    err = np.abs(rows - ref) / np.maximum(np.abs(ref), 1e-3 * scale[None, :])
    assert err.max() <= RECON_LOOSE, (int(np.argmax(err.max(axis=0))), float(err.max()))
    pol = rows[:, 53:61]
    assert np.all((pol[:, 1:] >= 0) & (pol[:, 1:] <= 2 * np.pi)) and np.all((pol[:, 0] >= 0) & (pol[:, 0] <= np.pi))
    # Sivers angle = phi_pq - phi_targ (mod 2 pi) for the reconstructed event (columns 33, 55, 57)
    d = np.mod(rows[:, 32] - pol[:, 1], 2 * np.pi)
    assert np.allclose(np.mod(d - pol[:, 3] + np.pi, 2 * np.pi) - np.pi, 0.0, atol=1e-9)
    # q is nearly along the beam and the polarisation at 80 degrees to it
    assert np.all(np.abs(pol[:, 0] - np.radians(80.0)) < np.radians(25.0))
    from simc_gfortran_b200.lib import ntuple_tags
    tags = ntuple_tags(cfg)
    assert len(tags) == 61 and tags[52] == "phipqi" and tags[53:61] == ["th_tarq", "phitarq", "beta", "phis", "phic", "betai", "phisi", "phici"]


def test_field_refusals(polcase):
    cfg, sim, orc = polcase
    s = Simc(cfg, mode="strict")              # no map set
    try:
        from simc_gfortran_b200 import load_optics_fixture
        for arm in (1, 5):
            s.set_optics(load_optics_fixture(arm))
        with pytest.raises(SimcError) as e:
            s.run(0, 100, 1, s.accum_clear())
        assert "field map" in str(e.value)
    finally:
        s.close()
