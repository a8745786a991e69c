"""CPU tests of the event-level oracle: analytic identities, the conditioning of
the reference's radiative weight, run-level sanity of the C1 configuration."""
import os

import numpy as np
import pytest

from simc_gfortran_b200 import config_from_deck, load_optics_fixture

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECK = os.path.join(ROOT, "decks", "c1_eep_hydrogen_hms_shms.inp")


@pytest.fixture(scope="module")
def cfg(built_lib):
    return config_from_deck(DECK)[0]


def clone(cfg):
    return type(cfg).from_buffer_copy(bytes(cfg))


def test_c1_config_matches_the_deck(cfg):
    assert cfg.doing_hyd_elast == 1 and cfg.doing_eep == 1 and cfg.electron_arm == 1 and cfg.hadron_arm == 5
    assert cfg.Mh == 938.27231 and cfg.targ.M == 938.27231
    assert abs(cfg.spec_e.theta - np.radians(25.90)) < 1e-15 and abs(cfg.spec_e.phi - 1.5 * np.pi) < 1e-15
    assert abs(cfg.targ.length - 0.295172 / 0.07332) < 1e-12                  # dbase.f:441-442
    assert cfg.using_Coulomb == 0                                              # forced off for Z=1, dbase.f:547
    assert abs(cfg.dEbeam - 8800 * 0.05 / 100) < 1e-12
    # SPedge widened by the HMS / SHMS slops (init.f:109-171, simulate.inc:20-44)
    assert abs(cfg.SPedge_e.delta.min + 10.5) < 1e-12 and abs(cfg.SPedge_p.xptar.max - 0.055) < 1e-12
    assert cfg.gen.e.yptar.max > cfg.SPedge_e.yptar.max                        # + extreme multiple scattering
    assert cfg.gen.sumEgen.min == 0.0 and cfg.gen.sumEgen.max == 0.0           # init.f:449-451
    assert cfg.Egamma1_max == cfg.Egamma_tot_max > 1000
    assert abs(cfg.etatzai - (12.0 + 2.0 / (5.31 + 6.144)) / 9.0) < 1e-15     # init.f:634
    assert 8799.0 < cfg.Ebeam_vertex_ave < 8800.0
    for k in range(8):
        assert cfg.hist_axis[0][k].bin > 0
    assert cfg.hist_axis[1][2].min == -cfg.gen.e.xptar.max                    # x' histograms hold -xptar


def test_elastic_identities_without_smearing(oracle_with_optics, cfg):
    """E' = E M/(M+E(1-cos)) (event.f:518): with radiation, energy loss and smearing off the
    reconstructed Em and Pm vanish up to the optics resolution."""
    c = clone(cfg)
    c.using_rad = 0; c.doing_tail[0] = c.doing_tail[1] = c.doing_tail[2] = 0
    c.using_Eloss = 0; c.correct_Eloss = 0; c.mc_smear = 0; c.dEbeam = 0.0
    c.Ebeam_vertex_ave = c.Ebeam
    rec, stage = oracle_with_optics.event_batch(c, 0, 20000, 11)
    ok = stage == 4
    assert ok.sum() > 1500
    assert np.abs(rec[44][ok]).mean() < 6.0 and np.abs(rec[45][ok]).mean() < 12.0     # Em, Pm in MeV at 8.8 GeV
    assert np.abs(rec[46][ok] - 938.27231).mean() < 6.0                                 # W
    assert np.all(rec[7][stage >= 1] == 1.0)                                            # gen_weight untouched
    # vertex kinematics obey the elastic relation exactly
    Ein, eE, Q2 = rec[10][stage >= 1], rec[11][stage >= 1], rec[19][stage >= 1]
    assert np.allclose(Q2, 2 * 938.27231 * (Ein - eE), rtol=1e-12)


def test_reference_radiative_weight_is_ill_conditioned(oracle_with_optics, cfg):
    """Evidence for the LOOSE tolerance of tests/test_loop_gpu.py: one ulp on Ebeam moves the
    oracle's own gen_weight by up to ~1e-10 relative (bremos forms ar1 ~ 1e-8 as 0.5 - 0.4999...,
    brem.f:411-418), while quantities upstream of the radiative constants move by ~1e-16."""
    n = 4000
    a, sa = oracle_with_optics.event_batch(cfg, 0, n, 1)
    c2 = clone(cfg)
    c2.Ebeam = np.nextafter(cfg.Ebeam, 1e9)
    b, sb = oracle_with_optics.event_batch(c2, 0, n, 1)
    m = (sa >= 1) & (sb >= 1)
    rel = lambda k: np.abs(a[k][m] - b[k][m]) / np.abs(a[k][m])
    assert rel(8).max() == 0.0                         # jacobian: angles only
    assert 1e-12 < rel(7).max() < 1e-7                 # gen_weight
    assert (rel(7) > 1e-12).mean() > 0.02


def test_c1_run_level_sanity(oracle_with_optics, cfg):
    acc = oracle_with_optics.run(cfg, 0, 30000, 4, threads=4)
    assert acc.ntried == 30000
    assert 0.15 < acc.nsuccess / acc.ntried < 0.27
    assert acc.npasscuts <= acc.ncontribute == acc.nsuccess
    avg_sig = acc.sum_sigcc.value() / acc.nsuccess
    assert 0.5 * cfg.w_ref < avg_sig < 2.0 * cfg.w_ref          # AVERAGE.sigcc vs CENTRAL.sigcc (simc.f:912-920)
    hn = np.ctypeslib.as_array(acc.hist_n)
    assert hn[2][1].sum() == 30000 and hn[1][0].sum() <= acc.nsuccess
    # the two threads-splittings give the same bits (integer accumulators)
    acc1 = oracle_with_optics.run(cfg, 0, 30000, 4, threads=1)
    assert bytes(acc) == bytes(acc1)
    # normalised yield = sum(weight) * luminosity * genvol / ntried is finite and positive
    genvol = (cfg.gen.e.yptar.max - cfg.gen.e.yptar.min) * (cfg.gen.e.xptar.max - cfg.gen.e.xptar.min)
    targetfac = cfg.targ.mass_amu / 3.75914e6 / (cfg.targ.abundancy / 100.) * abs(np.cos(cfg.targ.angle)) / (cfg.targ.thick * 1000.)
    lumi = 1.0 / targetfac
    y = acc.wtcontribute.value() * lumi * genvol / acc.ntried
    assert 5.0 < y < 80.0        # ~4e-6 ub/sr x 0.024 sr x 20% x 1.1e9 ub^-1/mC ~ 20 counts per mC


def test_counter_based_stream_is_statistically_equivalent_to_ranlux(oracle_with_optics, cfg):
    """The product replaces the reference's sequential RANLUX by a counter-based Philox stream.  The
    two cannot give the same events; they must give the same distributions: acceptance (binomial),
    normalised yield (weighted sum) and the generated / reconstructed count histograms (chi^2)."""
    n = 240000
    a = oracle_with_optics.run(cfg, 0, n, 11, threads=8, ranlux=True)
    b = oracle_with_optics.run(cfg, 0, n, 11, threads=8, ranlux=False)
    pa, pb = a.nsuccess / n, b.nsuccess / n
    sig = np.sqrt(pa * (1 - pa) / n + pb * (1 - pb) / n)
    assert abs(pa - pb) < 4.5 * sig, (pa, pb, sig)
    ya, yb = a.wtcontribute.value() / n, b.wtcontribute.value() / n
    # relative statistical error of a weighted yield ~ 1/sqrt(npasscuts) x (weight spread ~ 1.5)
    rel = 1.5 * np.sqrt(1.0 / a.npasscuts + 1.0 / b.npasscuts)
    assert abs(ya - yb) / yb < 4.5 * rel, (ya, yb, rel)
    ha, hb = np.ctypeslib.as_array(a.hist_n).astype(float), np.ctypeslib.as_array(b.hist_n).astype(float)
    for s_, k in ((2, 0), (2, 1), (2, 3), (1, 0), (1, 3), (1, 4)):          # geni / gen: e delta, e yptar, p delta, p yptar
        x, y = ha[s_, k], hb[s_, k]
        m = (x + y) > 20
        chi2 = ((x[m] - y[m]) ** 2 / (x[m] + y[m])).sum()
        ndf = m.sum()
        assert chi2 < ndf + 5 * np.sqrt(2 * ndf), (s_, k, chi2, ndf)
