"""Full runs, statistically: the GPU loop with its counter-based Philox stream against the CPU oracle driven by the
reference's own sequential RANLUX generator (cern/ranlux.f, one stream per thread), for the five BASELINE
configurations.  The two cannot produce the same events; they must produce the same physics: acceptance (binomial),
normalised yield (weighted sum), the 24 count histograms and the weighted RECON histograms (chi^2).

This is the comparison north_star asks for against a gfortran run, made with what exists here: the oracle restates the
Fortran routine by routine and takes its random numbers from the restated RANLUX (checked against James' published
vectors, tests/test_oracle_primitives.py).  No Fortran compiler exists in this image or on the GPU boxes, so
the reference side of the comparison is the restatement, not the binary.  For C1 the samples are large enough to hold
the normalised yields together at the 2e-3 level (3 sigma)."""
import os

import numpy as np
import pytest

from simc_gfortran_b200 import Simc, config_from_deck, load_optics_fixture
from tests.oracle_lib import load_cteq5_fixture, load_pfermi_fixture

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
#            deck                                arms     oracle tries  GPU tries
CASES = {
    "c1": ("c1_eep_hydrogen_hms_shms.inp", (1, 5), 12_000_000, 200_000_000),
    "c2": ("c2_eep_carbon_hms_sos.inp", (1, 2), 1_500_000, 40_000_000),
    "c3": ("c3_eepi_hydrogen_hms_shms.inp", (1, 5), 1_500_000, 40_000_000),
    "c4": ("c4_semi_deuterium_hms_shms.inp", (1, 5), 1_500_000, 40_000_000),
    "c5": ("c5_eek_hydrogen_hrsl_hrsr.inp", (4, 3), 3_000_000, 80_000_000),
}


def tables(cfg, sim, orc):
    if cfg.doing_heavy:
        z = np.load(os.path.join(ROOT, "simc_gfortran_b200", "data", "benharsf_12.npz"))
        sim.set_sf_table(z["pm"], z["em"], z["sf_proton"])
        orc.set_sf_table(z["pm"], z["em"], z["sf_proton"])
    if cfg.doing_semi:
        t = load_cteq5_fixture()
        sim.set_cteq5_table(t)
        orc.set_cteq5_table(t)
        pval, mprob = load_pfermi_fixture()
        sim.set_pfermi_table(pval, mprob)
        orc.set_pfermi_table(pval, mprob)


@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_run_agrees_with_the_ranlux_oracle(oracle_with_optics, name):
    deck, arms, n_cpu, n_gpu = CASES[name]
    cfg = config_from_deck(os.path.join(ROOT, "decks", deck))[0]
    orc = oracle_with_optics
    sim = Simc(cfg, mode="strict")
    try:
        for arm in arms:
            sim.set_optics(load_optics_fixture(arm))
        tables(cfg, sim, orc)
        threads = min(32, os.cpu_count() or 8)
        a = orc.run(cfg, 0, n_cpu, 20260, threads=threads, ranlux=True)
        b = sim.accum_clear()
        sim.run(0, n_gpu, 77, b)
    finally:
        sim.close()
    assert a.ntried == n_cpu and b.ntried == n_gpu and a.npasscuts > 2000
    # acceptance
    pa, pb = a.nsuccess / n_cpu, b.nsuccess / n_gpu
    sig = np.sqrt(pa * (1 - pa) / n_cpu + pb * (1 - pb) / n_gpu)
    assert abs(pa - pb) < 4.0 * sig, (name, pa, pb, sig)
    # normalised yield = wtcontribute * luminosity * genvol / ntried: the constants are the same on both sides.  The
    # statistical error of a weighted sum is sqrt(sum w^2); estimated from the RECON delta histogram's spread it is
    # ~ 1.5 / sqrt(npasscuts) for these weights (3 for the steep semi-inclusive and kaon cross sections)
    ya, yb = a.wtcontribute.value() / n_cpu, b.wtcontribute.value() / n_gpu
    spread = 1.5 if name in ("c1", "c2", "c3") else 3.0
    rel = spread * np.sqrt(1.0 / a.npasscuts + 1.0 / b.npasscuts)
    assert abs(ya - yb) / yb < 3.0 * rel, (name, ya, yb, rel)
    if name == "c1":
        assert 3.0 * rel < 3e-3, rel          # the C1 samples pin the yields to the few 1e-3 level
    # mean central cross section of the accepted events
    sa, sb = a.sum_sigcc.value() / a.nsuccess, b.sum_sigcc.value() / b.nsuccess
    assert abs(sa - sb) / sb < 3.0 * rel, (name, sa, sb)
    # count histograms: geni (all tries), gen (successes), RECON Em / Pm: shapes, scaled to the same number of entries
    ha, hb = np.ctypeslib.as_array(a.hist_n).astype(float), np.ctypeslib.as_array(b.hist_n).astype(float)
    checked = 0
    for s_ in range(3):
        for k in range(ha.shape[1]):
            x, y = ha[s_, k], hb[s_, k]
            if x.sum() < 1000 or y.sum() < 1000:
                continue
            f = x.sum() / y.sum()
            m = (x + y * f) > 40
            if m.sum() < 3:
                continue
            chi2 = ((x[m] - y[m] * f) ** 2 / (x[m] + y[m] * f * f)).sum()
            ndf = m.sum() - 1
            assert chi2 < ndf + 5 * np.sqrt(2 * ndf), (name, s_, k, chi2, ndf)
            checked += 1
    assert checked >= 12, checked
    # weighted RECON histograms (the six arm quantities): bin contents relative to the histogram's integral
    wa = np.array([[a.hist_w[k][i].value() for i in range(ha.shape[2])] for k in range(6)])
    wb = np.array([[b.hist_w[k][i].value() for i in range(ha.shape[2])] for k in range(6)])
    for k in range(6):
        x, y = wa[k] / wa[k].sum(), wb[k] / wb[k].sum()
        # effective entries of the smaller (CPU) sample: sum(w)^2 / sum(w^2); the spectral function (C2) and the steep
        # semi-inclusive / kaon cross sections spread the weights over more than a decade
        hs = {"c1": 1.5, "c2": 3.0, "c3": 2.0, "c4": 3.0, "c5": 3.0}[name]
        n_eff = a.nsuccess / hs ** 2
        err = np.sqrt(np.maximum(x, 1e-12) / n_eff)
        m = x > 5.0 / n_eff
        chi2 = (((x - y)[m] / err[m]) ** 2).sum()
        ndf = m.sum() - 1
        assert chi2 < ndf + 6 * np.sqrt(2 * ndf), (name, "RECON", k, chi2, ndf)
