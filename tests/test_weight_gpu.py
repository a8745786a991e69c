"""GPU parity test of the end of the loop body on identical dumped vectors (simc_b200_weight_batch):
complete_recon_ev (event.f:1056-1359), complete_main (event.f:1363-1569) with sigep / deForest + sf_lookup /
peepi / peeK / peepiX, pass_cuts and the hard cuts (simc.f:219-246).  The oracle dumps what those routines read for
real tries (after montecarlo); both sides then start from the same numbers.  Bar: success / pass_cuts flags
bit-exact, FP64 outputs within 1e-12 relative (natural scales for quantities that pass through zero)."""
import os

import numpy as np
import pytest

from simc_gfortran_b200 import Simc, config_from_deck, load_optics_fixture
from tests.oracle_lib import (Oracle, load_cteq5_fixture, load_pfermi_fixture)

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = 1e-12
# outputs: success, pass_cuts, weight, sigcc, sigcc_recon, Em, Pm, W, thetacm, phicm, sigcm, davejac, survivalprob, mm, wcm
NAMES = ["success", "pass_cuts", "weight", "sigcc", "sigcc_recon", "recon.Em", "recon.Pm", "recon.W", "thetacm", "phicm",
         "sigcm", "davejac", "survivalprob", "mm", "wcm"]


def sf_table():
    z = np.load(os.path.join(ROOT, "simc_gfortran_b200", "data", "benharsf_12.npz"))
    return z["pm"], z["em"], z["sf_proton"]


def setup(deck, edit=None):
    cfg = config_from_deck(os.path.join(ROOT, "decks", deck))[0]
    if edit:
        edit(cfg)
    orc = Oracle()
    sim = Simc(cfg, mode="strict")
    for arm in sorted({cfg.electron_arm, cfg.hadron_arm}):
        t = load_optics_fixture(arm)
        orc.set_optics(t)
        sim.set_optics(t)
    if cfg.doing_heavy:
        orc.set_sf_table(*sf_table())
        sim.set_sf_table(*sf_table())
    if cfg.doing_semi:
        orc.set_cteq5_table(load_cteq5_fixture())
        sim.set_cteq5_table(load_cteq5_fixture())
        orc.set_pfermi_table(*load_pfermi_fixture())
        sim.set_pfermi_table(*load_pfermi_fixture())
    return cfg, orc, sim


def compare(cfg, orc, sim, n_tries, seed, cols, min_rows):
    inp, valid = orc.weight_inputs(cfg, 0, n_tries, seed)
    inp = np.ascontiguousarray(inp[:, valid])
    assert inp.shape[1] >= min_rows, inp.shape
    ref = orc.weight_batch(cfg, inp)
    out = sim.weight_batch(inp)
    assert np.array_equal(out[0], ref[0]), "success flags differ"
    assert np.array_equal(out[1], ref[1]), "pass_cuts flags differ"
    ok = ref[0] == 1
    assert ok.sum() >= min_rows * 0.5
    worst = {}
    for k in cols:
        a, b = out[k][ok], ref[k][ok]
        assert np.array_equal(np.isnan(a), np.isnan(b)), NAMES[k]
        m = ~np.isnan(b)
        # Em, Pm of elastic events are differences of GeV-size numbers: scale 1 MeV; angles: 1e-3 rad
        scale = {5: 1.0, 6: 1.0, 8: 1e-3, 9: 1e-3, 13: 1.0}.get(k, 0.0)
        e = np.abs(a[m] - b[m]) / np.maximum(np.maximum(np.abs(a[m]), np.abs(b[m])), max(scale, 1e-300))
        worst[NAMES[k]] = float(e.max()) if e.size else 0.0
        assert e.size == 0 or e.max() <= RTOL, (NAMES[k], float(e.max()))
    return worst


def test_hydrogen_elastic():
    cfg, orc, sim = setup("c1_eep_hydrogen_hms_shms.inp")
    try:
        compare(cfg, orc, sim, 30000, 11, [2, 3, 4, 5, 6, 7], 3000)
    finally:
        sim.close()


@pytest.mark.parametrize("deforest_flag", [0, 1, -1])
def test_carbon_deforest_and_spectral_function(deforest_flag):
    """A(e,e'p): sf_lookup_diff on the Benhar table and all three off-shell prescriptions of deForest
    (physics_proton.f:23-135: sigma_cc1, sigma_cc2, on-shell)."""
    def edit(c):
        c.deForest_flag = deforest_flag
    cfg, orc, sim = setup("c2_eep_carbon_hms_sos.inp", edit)
    try:
        compare(cfg, orc, sim, 40000, 12, [2, 3, 4, 5, 6, 7], 1500)
    finally:
        sim.close()


def test_pion_electroproduction():
    cfg, orc, sim = setup("c3_eepi_hydrogen_hms_shms.inp")
    try:
        compare(cfg, orc, sim, 40000, 13, [2, 3, 5, 6, 7, 8, 9, 10, 11, 13, 14], 1500)
    finally:
        sim.close()


def test_kaon_electroproduction():
    cfg, orc, sim = setup("c5_eek_hydrogen_hrsl_hrsr.inp")
    try:
        compare(cfg, orc, sim, 150000, 14, [2, 3, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14], 800)
    finally:
        sim.close()


def test_semi_inclusive():
    cfg, orc, sim = setup("c4_semi_deuterium_hms_shms.inp")
    try:
        compare(cfg, orc, sim, 60000, 15, [2, 3, 5, 6, 7, 10, 11, 13], 1200)
    finally:
        sim.close()


def test_hard_cuts_change_success_not_weights():
    def edit(c):
        c.hard_cuts = 1
    cfg, orc, sim = setup("c1_eep_hydrogen_hms_shms.inp", edit)
    try:
        inp, valid = orc.weight_inputs(cfg, 0, 20000, 3)
        inp = np.ascontiguousarray(inp[:, valid])
        ref, out = orc.weight_batch(cfg, inp), sim.weight_batch(inp)
        assert np.array_equal(out[0], ref[0]) and np.array_equal(out[1], ref[1])
        assert (ref[0] == 0).sum() > 50 and np.array_equal(ref[0] == 1, (ref[1] == 1) & (ref[5] <= cfg.cuts_Em.max))
    finally:
        sim.close()
