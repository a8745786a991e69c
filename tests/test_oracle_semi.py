"""CPU tests of the oracle's semi-inclusive weight (semi_physics.f, cteq5/Ctq5Pdf.f, F1F2IN21_v1.0.f) and of
the host setup for C4.  The reference ships no known-answer values for these routines; what pins the
restatement here are identities the physics provides: the quark-number sum rules of the CTEQ5M table,
exactness of the three-point interpolation on quadratics, and closure between the two independent inputs of
peepiX -- the Christy fit's F2 in the deep-inelastic region against the parton-model F2 built from CTEQ5M."""
import os

import numpy as np
import pytest

from simc_gfortran_b200 import config_from_deck, load_optics_fixture
from tests.oracle_lib import load_cteq5_fixture, load_pfermi_fixture

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECK = os.path.join(ROOT, "decks", "c4_semi_deuterium_hms_shms.inp")


@pytest.fixture(scope="module")
def orc(oracle_with_optics):
    oracle_with_optics.set_cteq5_table(load_cteq5_fixture())
    oracle_with_optics.set_pfermi_table(*load_pfermi_fixture())
    return oracle_with_optics


def test_cteq5_fixture_shape():
    t = load_cteq5_fixture()
    assert (t["nx"], t["nt"], t["nfmx"]) == (90, 13, 5) and abs(t["lam"] - 0.226) < 1e-12
    assert len(t["upd"]) == 91 * 14 * 8 and t["xv"][0] == 0.0 and t["xv"][-1] == 1.0
    assert np.all(np.diff(t["xv"]) > 0) and np.all(np.diff(t["qv"]) > 0)


def test_cteq5_quark_number_sum_rules(orc):
    """int (u - ubar) dx = 2 and int (d - dbar) dx = 1 at any scale: checks the flavour offsets, the grid
    search and the interpolation of PartonX in one go."""
    # substitution x = t^3 concentrates points at small x where the valence distributions still vary fast
    t = (np.arange(200000) + 0.5) / 200000
    x = t ** 3
    w = 3 * t ** 2 / 200000
    for q in (1.5, 3.0, 10.0):
        qq = np.full_like(x, q)
        uv = orc.ctq5pdf_batch(1, x, qq) - orc.ctq5pdf_batch(-1, x, qq)
        dv = orc.ctq5pdf_batch(2, x, qq) - orc.ctq5pdf_batch(-2, x, qq)
        assert abs((uv * w).sum() - 2.0) < 0.03, (q, (uv * w).sum())
        assert abs((dv * w).sum() - 1.0) < 0.03, (q, (dv * w).sum())
        # strange sea is symmetric in this set: s = sbar (Ip = -Iprtn for |Iprtn| >= 3, Ctq5Pdf.f:170-174)
        assert np.array_equal(orc.ctq5pdf_batch(3, x[::100], qq[::100]), orc.ctq5pdf_batch(-3, x[::100], qq[::100]))


def test_cteq5_momentum_sum_rule(orc):
    t = (np.arange(100000) + 0.5) / 100000
    x = t ** 3
    w = 3 * t ** 2 / 100000
    qq = np.full_like(x, 2.0)
    tot = np.zeros_like(x)
    for ip in range(-5, 6):
        tot += orc.ctq5pdf_batch(ip, x, qq)
    assert abs((x * tot * w).sum() - 1.0) < 0.02


def test_cteq5_is_exact_at_grid_points(orc):
    t = load_cteq5_fixture()
    nx, nt = t["nx"], t["nt"]
    upd = t["upd"].reshape(8, nt + 1, nx + 1)
    ix = np.arange(5, nx - 1, 7)
    for iq in (2, 6, 11):
        for ip, jfl in ((1, 6), (-1, 4), (2, 7), (-2, 3), (3, 2), (0, 5)):
            v = orc.ctq5pdf_batch(ip, t["xv"][ix], np.full(len(ix), t["qv"][iq]))
            ref = np.maximum(upd[jfl, iq, ix], 0.0)
            assert np.allclose(v, ref, rtol=1e-9, atol=1e-12), (ip, iq)


def test_christy_fit_matches_parton_model_in_dis(orc):
    """F2p from the resonance fit (RESMODP) agrees with x sum e_q^2 (q + qbar) from CTEQ5M at W2 = 9-16 GeV2,
    Q2 = 4-8 GeV2 to the accuracy a leading-order comparison can have; F2n/F2p lies between 1/4 and 1."""
    w2 = np.array([9.0, 12.0, 16.0, 9.0, 12.0])
    q2 = np.array([4.0, 6.0, 8.0, 8.0, 4.0])
    f = orc.christy_batch(w2, q2)
    mp = 0.938272
    x = q2 / (q2 + w2 - mp * mp)
    q = np.sqrt(q2)
    pdf = {ip: orc.ctq5pdf_batch(ip, x, q) for ip in (1, -1, 2, -2, 3, -3)}
    f2_parton = x * (4. / 9 * (pdf[1] + pdf[-1]) + 1. / 9 * (pdf[2] + pdf[-2]) + 1. / 9 * (pdf[3] + pdf[-3]))
    assert np.all(np.abs(f[2] / f2_parton - 1.0) < 0.25), f[2] / f2_parton
    assert np.all((f[5] / f[2] > 0.25) & (f[5] / f[2] < 1.0)), f[5] / f[2]
    # Callan-Gross within the size of R = sigma_L / sigma_T: F2 ~ 2 x F1 (1 + R) / (1 + 4 M^2 x^2 / Q2)
    assert np.all(np.abs(f[2] / (2 * x * f[0]) - 1.0) < 0.35)


def test_christy_fit_vanishes_below_pion_threshold(orc):
    f = orc.christy_batch(np.array([1.10, 1.15, 1.17]), np.array([1.0, 2.0, 0.5]))      # (Mp + Mpi)^2 = 1.152
    assert np.all(f[:, 0] == 0.0) and np.all(f[:, 2:] > 0.0)


def test_delta_resonance_peak(orc):
    """The transverse strength peaks at the Delta(1232): W2 ~ 1.52 GeV2."""
    w2 = np.linspace(1.3, 1.9, 61)
    f = orc.christy_batch(w2, np.full_like(w2, 0.5))
    assert abs(w2[np.argmax(f[0])] - 1.5) < 0.05


def test_c4_config_from_deck():
    cfg, ngen, charge = config_from_deck(DECK)
    assert ngen == -1000000 and charge == 1.0
    assert cfg.doing_semi and cfg.doing_semipi and not cfg.doing_semika and not cfg.doing_pion     # dbase.f:125-129
    assert cfg.doing_deutsemi and not cfg.doing_hydsemi and not cfg.doing_hplus and not cfg.do_fermi
    assert not cfg.using_rad and cfg.doing_tail[0] == cfg.doing_tail[1] == cfg.doing_tail[2] == 0
    assert abs(cfg.Mh - 139.57018) < 1e-9
    assert abs(cfg.targ.M - 1875.613) < 1e-3 and cfg.targ.Mtar_struck == 938.27231 == cfg.targ.Mrec_struck
    # init.f:450-454: sumEgen.max = Ebeam_max + Mtar_struck - Mrec_struck; .min = edge.e.E.min + edge.p.E.min
    assert abs(cfg.gen.sumEgen.min - (cfg.edge.e.E.min + cfg.edge.p.E.min)) < 1e-9
    assert cfg.Egamma_tot_max == 0.0 and cfg.w_ref == 1e-9


def test_c4_loop_on_the_oracle(orc):
    cfg = config_from_deck(DECK)[0]
    rec, stage = orc.event_batch(cfg, 0, 6000, 2)
    done = stage == 4
    assert 50 < done.sum() < 1000 and (stage == 0).sum() > 100
    assert np.all(rec[6][done] >= 0) and 1e-10 < np.median(rec[6][done]) < 1e-7       # ub/MeV/sr^2
    rows, tries = orc.ntuple_batch(cfg, 0, 6000, 2)
    assert rows.shape == (done.sum(), 56)
    z_rec, z_vtx, pt2_vtx, x_vtx = rows[:, 43], rows[:, 44], rows[:, 46], rows[:, 48]
    assert np.all((z_vtx > 0.3) & (z_vtx <= 1.0)) and np.all(pt2_vtx >= 0) and np.all((x_vtx > 0.1) & (x_vtx < 0.8))
    assert np.median(np.abs(z_rec - z_vtx)) < 0.005 and np.abs(z_rec - z_vtx).max() < 0.3      # tails: pions that decayed
    # |p_fermi| column follows deut.dat: median of the deuteron momentum distribution ~ 45 MeV/c
    pf = np.abs(rows[:, 53]) * 1000.
    assert 25 < np.median(pf) < 70 and pf.max() < 1190.0
    # the weight is sigcc * jacobian * gen_weight (SF_weight = tgtweight = 1)
    assert np.allclose(rows[:, 40], rows[:, 38] * rec[8][tries] * rec[7][tries], rtol=1e-12)


def test_fermi_momentum_sampling_follows_table(orc):
    """event.f:337-353: invert the cumulative table.  With do_fermi the sampled x differs from Q2/2Mnu."""
    cfg = config_from_deck(DECK)[0]
    cfg.do_fermi = 1
    rows, _ = orc.ntuple_batch(cfg, 0, 20000, 8)
    pval, mprob = load_pfermi_fixture()
    pf = np.abs(rows[:, 53]) * 1000.
    # quartiles of the table
    for frac in (0.25, 0.5, 0.75):
        p_tab = pval[np.searchsorted(mprob / mprob[-1], frac)]
        # the acceptance and the fermi flux factor reweight only mildly
        assert abs(np.quantile(pf, frac) - p_tab) < 0.35 * p_tab + 5, (frac, np.quantile(pf, frac), p_tab)
    assert np.all(rows[:, 54] > 0) and np.abs(rows[:, 54] / rows[:, 48] - 1).max() > 0.02     # xfermi vs vertex xbj


def test_fdss_kaon_fragmentation_functions(orc):
    """fDSS (fdss/fdss.f): bilinear interpolation in (log z, log Q2) of z D(z) / ((1-z)^4 sqrt z); exact at grid
    nodes; favoured fragmentation (u -> K+, sbar -> K+) dominates; K- is the charge conjugate."""
    from tests.oracle_lib import load_fdss_fixture
    grid = load_fdss_fixture()
    orc.set_fdss_table(grid)
    xb = np.array([0.2, 0.4, 0.6, 0.8])
    ix = np.array([14, 22, 26, 30])
    q2 = np.full(4, 4.0)          # QS(5)
    kp = orc.fdss_batch(1, xb, q2)
    km = orc.fdss_batch(-1, xb, q2)
    utot, uval = grid[ix, 4, 0], grid[ix, 4, 6]
    assert np.allclose(kp[0], (utot + uval) / 2, rtol=1e-12) and np.allclose(kp[1], (utot - uval) / 2, rtol=1e-12)
    assert np.array_equal(kp[0], km[1]) and np.array_equal(kp[5], km[4])
    assert np.all(kp[0] > kp[1]) and np.all(kp[5] > kp[4])              # u -> K+ and sbar -> K+ are favoured
    mid = orc.fdss_batch(1, np.array([0.5]), np.array([np.sqrt(4.0 * 6.4)]))[0]
    lo, hi = orc.fdss_batch(1, np.array([0.5, 0.5]), np.array([4.0, 6.4]))[0]
    assert abs(mid - 0.5 * (lo + hi)) < 1e-12 * lo                       # linear in log Q2
