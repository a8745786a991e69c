"""CPU tests of the independent-particle spectral function (theory_init, init.f:828-905; the SF weight of
complete_main, event.f:1402-1428) and of the deck setup for D(e,e'p) and A(e,e'p) without use_benhar_sf.
Pins: each momentum distribution of the reference's theory files integrates to one nucleon after the
division by bs_norm, and the Lorentzian in Em integrates to one above E_Fermi."""
import os

import numpy as np
import pytest

from simc_gfortran_b200 import config_from_deck, load_optics_fixture
from tests.oracle_lib import load_theory_fixture, write_theory_file

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECKS = {"h2": "d1_eep_deuterium_hms_sos.inp", "c12": "c2t_eep_carbon_theory_hms_sos.inp"}


@pytest.fixture(scope="module")
def data_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("theory")
    for nm in ("h2", "c12"):
        write_theory_file(load_theory_fixture(nm), str(d / f"{nm}.theory"))
    return str(d)


def cfg_of(nm, data_dir):
    return config_from_deck(os.path.join(ROOT, "decks", DECKS[nm]), data_dir=data_dir)[0]


def test_deck_needs_the_theory_file():
    with pytest.raises(Exception) as ei:
        config_from_deck(os.path.join(ROOT, "decks", DECKS["h2"]))
    assert "h2.theory" in str(ei.value)


def test_deuterium_setup(data_dir):
    cfg = cfg_of("h2", data_dir)
    assert cfg.doing_deuterium and cfg.doing_eep and not cfg.doing_heavy and not cfg.doing_hyd_elast
    # init.f:326-330: Em is the binding energy, Pm up to the edge of the table (+-495 MeV/c)
    assert abs(cfg.VERTEXedge.Em.min - 2.22494) < 1e-3 and cfg.VERTEXedge.Em.min == cfg.VERTEXedge.Em.max
    assert cfg.VERTEXedge.Pm.max == 495.0
    # init.f:477-479, 500-503: the electron energy is generated over sumEgen, the proton energy over its acceptance
    assert cfg.gen.e.E.min == cfg.gen.sumEgen.min and cfg.gen.e.E.max == cfg.gen.sumEgen.max
    assert abs(cfg.targ.Mrec - 939.56563) < 1e-6


def test_theory_carbon_setup(data_dir):
    cfg = cfg_of("c12", data_dir)
    assert cfg.doing_heavy and not cfg.use_benhar_sf and not cfg.doing_deuterium
    assert cfg.VERTEXedge.Em.min >= 15.96 and cfg.VERTEXedge.Pm.max == 495.0      # E_Fermi of c12.theory


def test_momentum_distributions_are_normalised(oracle, data_dir):
    cfg = cfg_of("h2", data_dir)
    oracle.set_theory_table(load_theory_fixture("h2"), 0)
    p = np.linspace(0.0, 495.0, 49501)
    w = oracle.theory_batch(cfg, np.full_like(p, 2.2), p)
    n = np.trapezoid(4 * np.pi * p * p * w, p)
    assert abs(n - 1.0) < 0.01, n                       # one proton, absorption 1
    # linear interpolation between the tabulated points, exact at the points
    t = load_theory_fixture("h2")
    pk = t["pm_first"][0] + np.arange(60, 90) * t["pm_bin"][0]
    assert np.allclose(oracle.theory_batch(cfg, np.full_like(pk, 2.2), pk), t["rho"][60:90] / t["bs_norm"][0], rtol=1e-12)
    mid = oracle.theory_batch(cfg, np.array([2.2]), np.array([pk[3] + 5.0]))[0]
    assert abs(mid - 0.5 * (t["rho"][63] + t["rho"][64])) < 1e-12 * t["rho"][63]
    assert oracle.theory_batch(cfg, np.array([2.2]), np.array([600.0]))[0] == 0.0


def test_lorentzian_in_em_is_normalised_above_the_fermi_energy(oracle, data_dir):
    cfg = cfg_of("c12", data_dir)
    t = load_theory_fixture("c12")
    oracle.set_theory_table(t, 1)
    em = np.linspace(t["e_fermi"], 4000.0, 400001)
    p = np.linspace(0.0, 495.0, 100)
    tot = 0.0
    for pk in p[:99:33]:        # (at the table's edge the reference extrapolates: frac = 1.5)
        w = oracle.theory_batch(cfg, em, np.full_like(em, pk))
        # integral over Em of sum_i rho_i(p) nprot_i L_i(Em) = sum_i rho_i(p) nprot_i (tail beyond 4 GeV: < 0.4 %)
        w0 = 0.0
        pos = 0
        for m in range(2):
            n = t["n_pm"][m]
            pm = t["pm_first"][m] + np.arange(n) * t["pm_bin"][m]
            w0 += np.interp(pk, pm, t["rho"][pos:pos + n] / t["bs_norm"][m]) * t["nprot"][m] * t["absorption"]
            pos += n
        assert abs(np.trapezoid(w, em) / w0 - 1.0) < 0.01
    assert oracle.theory_batch(cfg, np.array([t["e_fermi"] - 0.1]), np.array([100.0]))[0] == 0.0


@pytest.mark.parametrize("nm", ["h2", "c12"])
def test_loop_on_the_oracle(nm, oracle_with_optics, data_dir):
    cfg = cfg_of(nm, data_dir)
    orc = oracle_with_optics
    orc.set_theory_table(load_theory_fixture(nm), cfg.doing_heavy)
    rec, stage = orc.event_batch(cfg, 0, 8000, 5)
    done = stage == 4
    assert done.sum() > 100
    assert np.all(rec[5][done] > 0)
    if nm == "h2":
        # reconstructed missing energy peaks at the deuteron binding energy (plus the radiative tail)
        assert abs(np.median(rec[44][done]) - 2.2) < 3.0
    else:
        assert np.median(rec[44][done]) > 15.0
