"""CPU checks of the oracle's reaction-specific physics (no GPU): identities and sanity ranges for
the meson-production weights (jacobians.f, physics_pion.f, physics_kaon.f) and the A(e,e'p)
spectral-function weight (sf_lookup.f, physics_proton.f deForest)."""
import os

import numpy as np
import pytest

from simc_gfortran_b200 import config_from_deck

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def deck(name):
    return config_from_deck(os.path.join(ROOT, "decks", name))[0]


def test_spectral_function_is_normalised(oracle):
    z = np.load(os.path.join(ROOT, "simc_gfortran_b200", "data", "benharsf_12.npz"))
    oracle.set_sf_table(z["pm"], z["em"], z["sf_proton"])
    pm, em = np.meshgrid(z["pm"], z["em"], indexing="ij")
    d = oracle.sf_batch(em.ravel(), pm.ravel())
    # sf_lookup_diff divides the bin content by 4 pi Pm^2 dPm dEm (sf_lookup.f:93): undo it on the grid
    total = (d * 4 * 3.1415926535 * pm.ravel() ** 2 * 5.0 * 20.0).sum()
    # sf_lookup_init normalises the table to one (sf_lookup.f:64-78).  Em exactly on the last grid point
    # matches no branch of sf_lookup.f:135-157 (`Em > Emval(numEm)` and `Em < Emval(iEm+1)` both fail), so
    # that column is missing from a sum over the grid itself.
    last = z["sf_proton"][:, -1].sum() / z["sf_proton"].sum()
    assert abs(total - (1.0 - last)) < 1e-9
    # beyond the last Pm bin the lookup sticks to the last column (w1 = 0, w2 = 1)
    a = oracle.sf_batch(np.array([22.5]), np.array([z["pm"][-1] + 100.0]))
    b = oracle.sf_batch(np.array([22.5]), np.array([z["pm"][-1]]))
    assert a[0] * (z["pm"][-1] + 100.0) ** 2 == pytest.approx(b[0] * z["pm"][-1] ** 2, rel=1e-12)


def test_carbon_eep_events(oracle_with_optics):
    z = np.load(os.path.join(ROOT, "simc_gfortran_b200", "data", "benharsf_12.npz"))
    oracle_with_optics.set_sf_table(z["pm"], z["em"], z["sf_proton"])
    cfg = deck("c2_eep_carbon_hms_sos.inp")
    assert cfg.doing_heavy and not cfg.doing_hyd_elast and cfg.VERTEXedge.Pm.max == 790.0
    rec, stage = oracle_with_optics.event_batch(cfg, 0, 20000, 11)
    done = stage == 4
    assert done.sum() > 500
    assert np.all(rec[5][done] > 0) and np.all(rec[6][done] > 0) and np.all(rec[9][done] > 0)   # weight, sigcc, sigcc_recon
    assert np.all(rec[45][done] < 900.0) and np.all(rec[44][done] > -50.0)                       # recon Pm, Em


@pytest.mark.parametrize("name,mrec", [("c5_eek_hydrogen_hrsl_hrsr.inp", 1115.68), ("c3_eepi_hydrogen_hms_shms.inp", 939.56563)])
def test_meson_production_events(oracle_with_optics, name, mrec):
    cfg = deck(name)
    rec, stage = oracle_with_optics.event_batch(cfg, 0, 30000, 5)
    done = stage == 4
    assert done.sum() > 100
    thetacm, phicm, sigcm, wcm, t = rec[48][done], rec[49][done], rec[50][done], rec[54][done], rec[55][done]
    assert np.all((thetacm >= 0) & (thetacm < 1.0)) and np.all((phicm >= 0) & (phicm <= 2 * np.pi + 1e-12))
    assert np.all(sigcm > 0) and np.all(rec[5][done] > 0)
    # W of the photon-nucleon system from the vertex: W^2 = Mp^2 + 2 Mp nu - Q2
    Mp = 938.27231
    nu = rec[10][done] - rec[11][done]
    assert np.allclose(wcm, np.sqrt(Mp * Mp + 2 * Mp * nu - rec[19][done]), rtol=1e-12)
    assert np.all(t > 0)                                              # -t in this convention (event.f:741)
    assert abs(np.median(rec[53][done]) - mrec) < 15.0                # missing mass = undetected baryon (+ radiative tail)
    surv = rec[52][done]
    if cfg.doing_kaon:
        assert np.all((surv > 0.03) & (surv < 0.5))                   # ~25 m of flight at beta*gamma*c*tau ~ 9.7 m
    else:
        assert np.all(surv == 1.0)


def test_delta_production_events(oracle_with_optics):
    """H(e,e'p)pi0 (physics_delta.f): the missing mass of the detected e + p is the pion's, the weight is flux times
    the CM jacobian (no model factor), and sig_blok's value rides along in the sigcm column."""
    cfg = deck("e1_eep_pi0_hydrogen_sos_hms.inp")
    assert cfg.doing_delta and not cfg.doing_pion and cfg.using_rad == 0
    rec, stage = oracle_with_optics.event_batch(cfg, 0, 40000, 5)
    done = stage == 4
    assert done.sum() > 300 and (stage == 0).sum() > 1000             # the low root of the quadratic mostly fails the limits
    assert abs(np.median(rec[53][done]) - 139.57018) < 25.0           # missing mass ~ m_pi (resolution of a 5 GeV/c proton)
    wcm, sigcm, davejac, sigcc = rec[54][done], rec[50][done], rec[51][done], rec[6][done]
    assert np.all((wcm > 938.27231 + 139.57018) & (wcm < 2200.0))     # above the pion threshold; the SOS window is wide
    assert np.all(np.isfinite(sigcm)) and np.all(sigcc > 0)           # the Brauel fit may go negative out here; it is not in the weight
    # sigcc = jacobian * gtpr (pfer = 0 -> fac = 1): recompute the flux from the vertex columns
    Mp, alpha = 938.27231, 1.0 / 137.0359895
    Ein, eE, Q2 = rec[10][done], rec[11][done], rec[19][done]
    k_eq = (wcm * wcm - Mp * Mp) / 2.0 / Mp
    nu = Ein - eE
    th2 = Q2 / (4.0 * Ein * eE - Q2)                                   # tan^2(theta/2) from Q2 = 4 E E' sin^2(theta/2)
    eps = 1.0 / (1.0 + 2.0 * (1.0 + nu * nu / Q2) * th2)
    gtpr = alpha / 2.0 / np.pi ** 2 * eE / Ein * k_eq / Q2 / (1.0 - eps)
    coul = (1.0 + cfg.targ.Coulomb_ave / cfg.Ebeam) ** 2 if cfg.using_Coulomb else 1.0
    assert np.allclose(sigcc, davejac * gtpr * coul, rtol=1e-9)


def test_sigmaid_nearest_bin_lookup(oracle):
    """sigmaid (physics_pion.f:577-728): sig0 = ST (1 + eps L/T + ...) from the nearest (Q2, W, cos theta*) bin;
    zero below W = 1.08 GeV and for unphysical kinematics; on the Delta peak sigma_T is several ub/sr."""
    from tests.oracle_lib import load_maid_fixture
    tbl = load_maid_fixture(3)
    oracle.set_maid_table(3, tbl)
    try:
        one = np.ones(1)
        assert oracle.sigmaid_batch(3, 0.4 * one, 1.07 * one, 1.645 * one, 0.5 * one, 0.3 * one)[0] == 0.0
        assert oracle.sigmaid_batch(3, 0.4 * one, 1.9 * one, 1.0 * one, 0.5 * one, 0.3 * one)[0] == 0.0     # nu > E0
        q2, w, e0, cth, phi = 0.4, 1.232, 1.645, 0.5, np.pi / 2
        am = 0.9383
        nu = (w * w - am * am + q2) / 2 / am
        sin2 = q2 / 4 / e0 / (e0 - nu)
        eps = 1 / (1 + 2 * (1 + nu * nu / q2) * sin2 / (1 - sin2))
        row = tbl[int((q2 + 0.1) / 0.2) - 1, int((w - 1.090) / 0.020) - 1, 2]          # cos theta* in [0.44, 0.63]
        st = row[0] / max(0.2, q2)
        want = st * (1 + eps * row[1] + np.sqrt(2 * eps * (1 + eps)) * np.cos(phi) * row[2] + eps * np.cos(2 * phi) * row[3])
        got = oracle.sigmaid_batch(3, q2 * one, w * one, e0 * one, cth * one, phi * one)[0]
        assert abs(got - want) < 1e-12 * abs(want) and 2.0 < got < 40.0
    finally:
        oracle.set_maid_table(3, None)


def test_generate_em_follows_the_spectral_function(oracle):
    """generate_em (sf_lookup.f:181-245): Em is drawn from the table's Em distribution at the given Pm -- the
    fraction of draws in the first bin equals the normalised strength of that bin."""
    from tests.oracle_lib import load_he3_fixtures
    _, sf = load_he3_fixtures()
    oracle.set_sf_table(sf["pm"], sf["em"], sf["sf_proton"])
    oracle.set_sf_em_widths(sf["dem"])
    ipm = 4
    pm = float(sf["pm"][ipm])
    em = oracle.generate_em_batch(11, np.full(40000, pm))
    col = sf["sf_proton"][ipm].copy()
    col[-1] = col[-2] + (col[-1] - col[-2])          # last grid point: the reused interval gives the grid value
    frac = col / col.sum()
    first = (np.abs(em - sf["em"][0]) <= sf["dem"][0] / 2).mean()
    assert abs(first - frac[0]) < 4 * np.sqrt(frac[0] * (1 - frac[0]) / len(em)) + 1e-3
    assert em.min() >= sf["em"][0] - sf["dem"][0] / 2 and em.max() <= sf["em"][-1] + sf["dem"][-1] / 2


def test_rho_production(oracle_with_optics):
    """H(e,e'rho0): generate_rho.f, rho_decay.f, rho_physics.f restated.  The Breit-Wigner mass (a tangent transform of one
    uniform number, cut at +-500 MeV), the decay-angle bookkeeping and the kinematics of the detected pion."""
    cfg = deck("r1_eerho_hydrogen_sos_hms.inp")
    assert cfg.doing_rho and abs(cfg.Mh - 769.3) < 1e-12
    n = 200000
    rec, stage = oracle_with_optics.event_batch(cfg, 0, n, 11)
    m = rec[58]
    thrown = m != 0.0                     # every try that reached generate_rho's first draw
    assert thrown.mean() > 0.95
    m = m[thrown]
    assert np.all(np.abs(m - 769.3) <= 500.0)
    # non-relativistic Breit-Wigner of width 150.2 MeV: P(|m - M| < Gamma/2) = atan(1) / atan(2*500/150.2)
    inside = (np.abs(m - 769.3) < 75.1).mean()
    expect = np.arctan(1.0) / np.arctan(2 * 500.0 / 150.2)
    assert abs(inside - expect) < 4 * np.sqrt(expect * (1 - expect) / len(m))
    assert abs(np.median(m) - 769.3) < 1.0
    gen = stage >= 1
    assert gen.sum() > 2000 and (stage == 4).sum() > 30
    # after rho_decay: the pion points into the hadron arm's hemisphere, its energy is below the rho's
    assert np.all(rec[21][gen] < rec[15][gen])                       # orig.p.E (pion) < vertex.p.E (rho)
    assert np.all((rec[59][gen] >= 0) & (rec[59][gen] <= np.pi))
    # hydrogen: no jacobian for the hadron angles, the electron's only (event.f:1013-1023)
    r = np.sqrt(1 + rec[13][gen] ** 2 + rec[14][gen] ** 2)
    assert np.allclose(rec[8][gen], 1 / r ** 3, rtol=1e-13)
    # the weight: peerho is positive, falls with -t (steep exponential slope)
    done = stage == 4
    assert np.all(rec[6][done] > 0) and np.all(rec[6][done] < 1e-4)
    rows, tries = oracle_with_optics.ntuple_batch(cfg, 0, n, 11)
    assert rows.shape[1] == 59 and len(rows) == done.sum()
    # p(e,e'pi)X with X = p + pi: the missing mass starts at Mp + Mpi; nucleus = nucleon for hydrogen
    assert np.all(rows[:, 33] > 0.93827 + 0.13957 - 0.02)
    assert np.allclose(rows[:, 58], rows[:, 33], rtol=1e-12)


def test_pizero_into_calorimeter(oracle_with_optics):
    """H(e,e'pi0)p with the NPS as the hadron arm: pizero_decay.f and calo/mc_calo.f restated.  The photons carry the pi0
    four-momentum, a kept event has pizero_ngamma photons inside the calorimeter face, and the hit is the straight line
    from the vertex over drift_to_cal."""
    cfg = deck("z1_eepi0_hydrogen_hms_nps.inp")
    assert cfg.doing_pizero and cfg.hadron_arm == 8 and abs(cfg.Mh - 134.9766) < 1e-12
    n = 40000
    rows, tries = oracle_with_optics.ntuple_batch(cfg, 0, n, 3)
    assert rows.shape[1] == 65 and len(rows) > 80
    g1, g2 = rows[:, 55:59], rows[:, 61:65]
    tot = g1 + g2
    assert np.allclose(np.sqrt(tot[:, 0] ** 2 - (tot[:, 1:] ** 2).sum(axis=1)), 134.9766, rtol=1e-9)
    assert np.all(np.abs(rows[:, [53, 59]]) <= 36.9) and np.all(np.abs(rows[:, [54, 60]]) <= 30.75)
    # photon 1 in the calorimeter frame (arm 8: rotation by +theta about x, simc.f:1496-1503), drifted 300 cm from the
    # vertex position in that frame: ycal - y0 = 300 * ey/ez with |y0| of a few cm (10 cm target seen at 7.5 degrees)
    th = cfg.spec_p.theta
    ey = g1[:, 2] * np.cos(th) - g1[:, 3] * np.sin(th)
    ez = g1[:, 2] * np.sin(th) + g1[:, 3] * np.cos(th)
    assert np.all(np.abs(rows[:, 54] - 300.0 * ey / ez) < 2.0)
    assert np.all(np.abs(rows[:, 53] - 300.0 * g1[:, 1] / ez) < 1.0)
    one = type(cfg).from_buffer_copy(bytes(cfg))
    one.pizero_ngamma = 1
    rows1, _ = oracle_with_optics.ntuple_batch(one, 0, n, 3)
    assert len(rows1) > len(rows)                     # one photon is enough: more events, some with a missed photon
    assert ((rows1[:, 53] == -1.0e10) | (rows1[:, 59] == -1.0e10)).any()
