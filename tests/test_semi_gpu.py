"""GPU parity tests for semi-inclusive production (C4 of BASELINE.json): D(e,e'pi-)X and H(e,e'pi+)X with
the CTEQ5M parton distributions, Bosted's fragmentation fit and the Christy 2021 inclusive fit.
Stage level: peepiX on dumped vertex vectors within 1e-12; loop level: per-try records, exact accumulators
and ntuple rows of libsimc_b200 against the CPU oracle on the same counter-based random stream."""
import os

import numpy as np
import pytest

from simc_gfortran_b200 import Simc, config_from_deck, load_optics_fixture
from tests.oracle_lib import load_cteq5_fixture, load_pfermi_fixture
from tests.test_loop_gpu import LOOSE, RECON_LOOSE, RTOL, SCALE, accum_equal_exact, rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECK = os.path.join(ROOT, "decks", "c4_semi_deuterium_hms_shms.inp")
SC = SCALE.copy()
SC[5] = SC[6] = 1e-12      # weights and cross sections of a few 1e-9 ub/MeV/sr^2
SC[50] = 1e-3              # ntup.sigcm = sighad
SC[51] = 1.0               # davejac
SC[53] = 1.0               # missing mass, MeV
SC[55] = 1e3               # t, MeV^2


def _variant(cfg, name):
    """The C4 deck and three neighbours: pi+; hydrogen target; deuterium with Fermi motion used and decay off."""
    if name == "d_piplus":
        cfg.doing_hplus = 1
    elif name == "d_fermi_nodecay":
        cfg.do_fermi = 1
        cfg.doing_decay = 0
    return cfg


@pytest.fixture(scope="module", params=["d_piminus", "d_piplus", "d_fermi_nodecay"])
def case(request, oracle_with_optics):
    cfg = _variant(config_from_deck(DECK)[0], request.param)
    orc = oracle_with_optics
    t, (pval, mprob) = load_cteq5_fixture(), load_pfermi_fixture()
    orc.set_cteq5_table(t)
    orc.set_pfermi_table(pval, mprob)
    s = Simc(cfg, mode="strict")
    for arm in (1, 5):
        s.set_optics(load_optics_fixture(arm))
    s.set_cteq5_table(t)
    s.set_pfermi_table(pval, mprob)
    yield request.param, cfg, s, orc
    s.close()


def semi_inputs(n, seed, fermi):
    """Vertex vectors spread over the C4 acceptance (and a bit beyond: thresholds, x -> 1)."""
    r = np.random.default_rng(seed)
    Ein = 5492.04 + r.uniform(-3, 3, n)
    eE = r.uniform(1300., 1950., n)
    th = np.radians(r.uniform(25., 32., n))
    nu = Ein - eE
    Q2 = 2 * Ein * eE * (1 - np.cos(th))
    q = np.sqrt(Q2 + nu * nu)
    uq = r.normal(size=(3, n)); uq /= np.linalg.norm(uq, axis=0)
    zhad = r.uniform(0.2, 1.0, n)
    pt2 = r.uniform(0., 2.5e5, n)
    thpq = np.where(r.uniform(size=n) < 0.5, 0.0, r.uniform(0, 0.1, n))
    pfer = r.uniform(0., 400., n) if fermi else np.zeros(n)
    pu = r.normal(size=(3, n)); pu /= np.linalg.norm(pu, axis=0)
    efer = 1875.613 - np.sqrt(939.56563 ** 2 + pfer ** 2) if fermi else np.full(n, 938.27231)
    return np.vstack([Ein, eE, nu, Q2, q, uq, pt2, zhad, thpq, pfer, pu, efer])


def test_peepix_on_dumped_vectors(case):
    """north_star: weights within 1e-12 relative on identical dumped per-event input vectors."""
    name, cfg, sim, orc = case
    inp = semi_inputs(20000, 5, fermi=bool(cfg.do_fermi))
    ref = orc.semi_batch(cfg, inp)
    out = sim.semi_batch(inp)
    assert (ref[0] > 0).sum() > 15000 and (ref[0] == 0).sum() > 5          # both sides of the 2-pion threshold
    assert np.array_equal(out[0] == 0, ref[0] == 0)
    assert not out[15].any()
    scale = [1e-14, 1e-6, 1e-3, 1e-3] + [1e-6] * 6 + [1e-6] * 4 + [1e-14]
    names = ["sigma_eepiX", "sighad", "davejac", "xbj", "u", "ubar", "d", "dbar", "s", "sbar", "F1p", "F2p", "F1n",
             "F2n", "sige"]
    for k, nm in enumerate(names):
        # NaN where the reference's own formulas give NaN (e.g. x clipped to 1: all parton densities vanish, 0/0)
        assert np.array_equal(np.isnan(out[k]), np.isnan(ref[k])), nm
        e = rel_err(out[k], ref[k], scale[k])
        assert np.nanmax(e) <= RTOL, (nm, float(np.nanmax(e)))


def test_event_records(case):
    name, cfg, sim, orc = case
    n = 40000
    ref, ref_stage = orc.event_batch(cfg, 500, n, 31)
    rec, stage = sim.event_batch(500, n, 31)
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
        ([1, 9, 44, 45, 46, 52, 53], stage == 4, RECON_LOOSE),
        ([5, 6, 48, 49, 50, 51, 54, 55], stage == 4, LOOSE),
    )
    for fields, mask, tol in groups:
        for k in fields:
            assert np.array_equal(np.isnan(rec[k][mask]), np.isnan(ref[k][mask])), names[k]
            e = rel_err(rec[k][mask], ref[k][mask], SC[k])
            assert np.nanmax(e) <= tol, (names[k], float(np.nanmax(e)))
    done = stage == 4
    assert done.sum() > 100 and (stage == 0).sum() > 10
    assert np.all(rec[6][done][~np.isnan(rec[6][done])] >= 0) and np.nanmedian(rec[6][done]) > 1e-10
    assert np.isnan(rec[6][done]).sum() <= 5          # do_fermi: x > 1 is clipped to 1 where all densities vanish (0/0)
    if name == "d_fermi_nodecay":
        assert np.all((rec[52][done] > 0.3) & (rec[52][done] < 1.0))       # survival to the back of the SHMS hut
    else:
        assert np.all(rec[52][done] == 1.0)


def test_accumulators_against_oracle(case):
    name, cfg, sim, orc = case
    n = 60000
    ref = orc.run(cfg, 0, n, 4, threads=8)
    acc = sim.accum_clear()
    sim.run(0, n, 4, acc)
    accum_equal_exact(acc, ref)
    # do_fermi: a hard nucleon along q can make x negative -- a `Stop` in the reference's Ctq5Pdf; counted here
    assert acc.unsupported == ref.unsupported and acc.unsupported <= (3 if name == "d_fermi_nodecay" else 0)
    assert acc.nsuccess > 200
    for f in ("wtcontribute", "sum_sigcc"):
        a, b = getattr(acc, f).value(), getattr(ref, f).value()
        assert abs(a - b) <= RECON_LOOSE * abs(b), f
        assert getattr(acc, f).qexp == getattr(ref, f).qexp
    for k in range(30):
        assert abs(acc.contrib[k].lo - ref.contrib[k].lo) <= 1e-7 * max(1.0, abs(ref.contrib[k].lo)), k
        assert abs(acc.contrib[k].hi - ref.contrib[k].hi) <= 1e-7 * max(1.0, abs(ref.contrib[k].hi)), k


def test_ntuple_rows(case):
    name, cfg, sim, orc = case
    n = 30000
    ref, ref_try = orc.ntuple_batch(cfg, 100, n, 12)
    rows, tries = sim.ntuple_batch(100, n, 12)
    assert rows.shape == ref.shape and rows.shape[1] == 56 and rows.shape[0] > 100
    assert np.array_equal(tries, ref_try)
    scale = np.ones(56)
    scale[38] = scale[40] = 1e-12      # sigcc, weight
    scale[50] = 1e-3
    for k in range(56):
        e = rel_err(rows[:, k], ref[:, k], scale[k])
        assert np.nanmax(e) <= RECON_LOOSE, (k, float(np.nanmax(e)))
        assert np.array_equal(np.isnan(rows[:, k]), np.isnan(ref[:, k])), k


def test_batching_independence(case):
    name, cfg, sim, orc = case
    a = sim.accum_clear()
    sim.set_batch(1 << 20)
    sim.run(0, 30000, 9, a)
    b = sim.accum_clear()
    sim.set_batch(2048)
    sim.run(0, 11000, 9, b)
    sim.run(11000, 19000, 9, b)
    sim.set_batch(1 << 20)
    assert bytes(a) == bytes(b)


def test_hydrogen_target(oracle_with_optics):
    """H(e,e'pi+)X: no momentum distribution needed, no Fermi draws; the ntuple's p_fermi column is 0/0 = NaN
    in the reference (results_write.f:211-212) and here."""
    cfg = config_from_deck(DECK)[0]
    cfg.targ.A = 1.0; cfg.targ.N = 0.0; cfg.targ.M = 938.27231; cfg.targ.Mrec = 0.0
    cfg.doing_deutsemi = 0; cfg.doing_hydsemi = 1; cfg.doing_hplus = 1
    orc = oracle_with_optics
    t = load_cteq5_fixture()
    orc.set_cteq5_table(t)
    s = Simc(cfg, mode="strict")
    try:
        for arm in (1, 5):
            s.set_optics(load_optics_fixture(arm))
        s.set_cteq5_table(t)
        ref = orc.run(cfg, 0, 40000, 6, threads=8)
        acc = s.accum_clear()
        s.run(0, 40000, 6, acc)
        accum_equal_exact(acc, ref)
        assert acc.nsuccess > 200
        a, b = acc.wtcontribute.value(), ref.wtcontribute.value()
        assert abs(a - b) <= RECON_LOOSE * abs(b)
        rows, _ = s.ntuple_batch(0, 20000, 6)
        rref, _ = orc.ntuple_batch(cfg, 0, 20000, 6)
        assert rows.shape == rref.shape and np.isnan(rows[:, 53]).all() and np.isnan(rref[:, 53]).all()
    finally:
        s.close()


def test_tables_are_required():
    cfg = config_from_deck(DECK)[0]
    s = Simc(cfg, mode="strict")
    try:
        for arm in (1, 5):
            s.set_optics(load_optics_fixture(arm))
        acc = s.accum_clear()
        with pytest.raises(Exception) as ei:
            s.run(0, 100, 1, acc)
        assert "CTEQ5" in str(ei.value)
        s.set_cteq5_table(load_cteq5_fixture())
        with pytest.raises(Exception) as ei:
            s.run(0, 100, 1, acc)
        assert "momentum distribution" in str(ei.value)
    finally:
        s.close()


# ---- semi-inclusive kaons: DSS fragmentation functions (fdss/fdss.f) ------------------------------------
@pytest.fixture(scope="module", params=["d_kplus_nodecay", "d_kminus"])
def kaon_case(request, oracle_with_optics, tmp_path_factory):
    from tests.oracle_lib import load_fdss_fixture, write_fdss_file
    txt = open(DECK).read()
    txt = txt.replace("doing_kaon = 0", "doing_kaon = 1").replace("doing_pion = 1", "doing_pion = 0").replace("ctau = 780.4", "ctau = 371.3")
    if request.param == "d_kplus_nodecay":
        txt = txt.replace("doing_hplus = 0", "doing_hplus = 1").replace("doing_decay = 1", "doing_decay = 0")
    d = tmp_path_factory.mktemp("semika")
    path = str(d / "deck.inp")
    open(path, "w").write(txt)
    cfg = config_from_deck(path)[0]
    assert cfg.doing_semika and not cfg.doing_semipi and not cfg.doing_kaon and abs(cfg.Mh - 493.677) < 1e-9
    orc = oracle_with_optics
    t, (pval, mprob), grid = load_cteq5_fixture(), load_pfermi_fixture(), load_fdss_fixture()
    orc.set_cteq5_table(t)
    orc.set_pfermi_table(pval, mprob)
    orc.set_fdss_table(grid)
    s = Simc(cfg, mode="strict")
    for arm in (1, 5):
        s.set_optics(load_optics_fixture(arm))
    s.set_cteq5_table(t)
    s.set_pfermi_table(pval, mprob)
    gpath = str(d / "KANLO.GRID")
    write_fdss_file(grid, gpath)
    s.load_fdss_file(gpath)
    yield request.param, cfg, s, orc
    s.close()


def test_kaon_peepix_on_dumped_vectors(kaon_case):
    name, cfg, sim, orc = kaon_case
    inp = semi_inputs(20000, 6, fermi=False)
    ref = orc.semi_batch(cfg, inp)
    out = sim.semi_batch(inp)
    assert (ref[0] > 0).sum() > 10000
    assert np.array_equal(out[0] == 0, ref[0] == 0)
    scale = [1e-15, 1e-7, 1e-3, 1e-3] + [1e-6] * 10 + [1e-14]
    for k in range(15):
        assert np.array_equal(np.isnan(out[k]), np.isnan(ref[k])), k
        e = rel_err(out[k], ref[k], scale[k])
        assert np.nanmax(e) <= RTOL, (k, float(np.nanmax(e)))


def test_kaon_loop(kaon_case):
    name, cfg, sim, orc = kaon_case
    n = 60000
    ref = orc.run(cfg, 0, n, 4, threads=8)
    acc = sim.accum_clear()
    sim.run(0, n, 4, acc)
    accum_equal_exact(acc, ref)
    assert acc.unsupported == ref.unsupported == 0 and acc.nsuccess > 200
    a, b = acc.wtcontribute.value(), ref.wtcontribute.value()
    assert abs(a - b) <= RECON_LOOSE * abs(b)
    rec, stage = sim.event_batch(0, 20000, 8)
    rref, sref = orc.event_batch(cfg, 0, 20000, 8)
    assert np.array_equal(stage, sref)
    done = stage == 4
    for k in (5, 6, 50, 51, 52):
        e = rel_err(rec[k][done], rref[k][done], SC[k])
        assert e.max() <= RECON_LOOSE, (k, float(e.max()))
    if name == "d_kplus_nodecay":
        # kaons over ~20 m of SHMS at 3.2 GeV/c (c tau = 3.7 m): survival applied to the weight (event.f:1560)
        # (events below the two-hadron threshold leave peepiX before the survival probability: weight 0, prob 1)
        live = done & (rec[6] > 0)
        assert live.sum() > 50 and np.all((rec[52][live] > 0.2) & (rec[52][live] < 0.7))
        assert np.all(rec[52][done & (rec[6] == 0)] == 1.0)
    else:
        assert np.all(rec[52][done] == 1.0)


def test_kaon_needs_the_grid(kaon_case):
    name, cfg, sim, orc = kaon_case
    s2 = Simc(cfg, mode="strict")
    try:
        for arm in (1, 5):
            s2.set_optics(load_optics_fixture(arm))
        s2.set_cteq5_table(load_cteq5_fixture())
        s2.set_pfermi_table(*load_pfermi_fixture())
        acc = s2.accum_clear()
        with pytest.raises(Exception) as ei:
            s2.run(0, 100, 1, acc)
        assert "DSS" in str(ei.value)
    finally:
        s2.close()
