"""GPU parity of the whole loop for the option branches the shipped decks do not take:
raster patterns 2 and 3 (event.f:160-200), target cans 2 and 3 (target.f:1-306), one_tail = -3 / 1 / 2 (radc_init,
init.f:576-651: which tails radiate), correct_raster off (simc.f:1441,1463: the SOS / HRS reconstruction then gets no
raster position), hard_cuts (simc.f:243-246), correct_Eloss off.  Each variant is a deck edit; both sides read the
edited deck through the product's deck reader, run the same tries of the same counter-based stream, and must agree
exactly on every counter, STOP counter and count histogram and to LOOSE on the weight sum."""
import os
import re

import numpy as np
import pytest

from simc_gfortran_b200 import Simc, config_from_deck, load_optics_fixture
from tests.test_loop_gpu import LOOSE, accum_equal_exact

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

VARIANTS = {
    # name: (deck, {key: value})
    "raster_circle": ("c1_eep_hydrogen_hms_shms.inp", {"targ%fr_pattern": "2", "targ%fr1": "0.05", "targ%fr2": "0.2"}),
    "raster_flat": ("c1_eep_hydrogen_hms_shms.inp", {"targ%fr_pattern": "3", "targ%fr1": "0.15", "targ%fr2": "0.1"}),
    "can_pudding": ("c1_eep_hydrogen_hms_shms.inp", {"targ%can": "2"}),
    "can_cryo2017": ("c1_eep_hydrogen_hms_shms.inp", {"targ%can": "3"}),
    "no_proton_tail": ("c1_eep_hydrogen_hms_shms.inp", {"one_tail": "-3"}),
    "tail1_only": ("c1_eep_hydrogen_hms_shms.inp", {"one_tail": "1"}),
    "tail2_only": ("c1_eep_hydrogen_hms_shms.inp", {"one_tail": "2"}),
    "hard_cuts": ("c1_eep_hydrogen_hms_shms.inp", {"hard_cuts": "1", "cuts%Em%max": "40."}),
    "no_eloss_correction": ("c1_eep_hydrogen_hms_shms.inp", {"correct_Eloss": "0"}),
    "hrs_no_raster_correction": ("c5_eek_hydrogen_hrsl_hrsr.inp", {"correct_raster": "0", "targ%fr1": "0.2", "targ%fr2": "0.2"}),
    "hrs_raster_correction": ("c5_eek_hydrogen_hrsl_hrsr.inp", {"correct_raster": "1", "targ%fr1": "0.2", "targ%fr2": "0.2"}),
    "sos_raster_circle": ("c2_eep_carbon_hms_sos.inp", {"targ%fr_pattern": "2", "targ%fr1": "0.0", "targ%fr2": "0.15", "correct_raster": "1"}),
}


def edited_deck(tmp_path, name):
    deck, edits = VARIANTS[name]
    text = open(os.path.join(ROOT, "decks", deck)).read()
    for key, val in edits.items():
        pat = re.compile(r"^(\s*" + re.escape(key) + r"\s*=\s*)([^;\n]*)", re.M)
        assert pat.search(text), f"{key} not in {deck}"
        text = pat.sub(lambda m: m.group(1) + val + "\t", text, count=1)
    path = tmp_path / f"{name}.inp"
    path.write_text(text)
    return str(path)


@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_variant(tmp_path, name, oracle_with_optics):
    orc = oracle_with_optics
    cfg = config_from_deck(edited_deck(tmp_path, name))[0]
    for key, val in VARIANTS[name][1].items():          # the deck reader took the edit
        if key == "targ%fr_pattern":
            assert cfg.targ.fr_pattern == int(val)
        if key == "targ%can":
            assert cfg.targ.can == int(val)
        if key == "correct_raster":
            assert cfg.correct_raster == int(val)
        if key == "hard_cuts":
            assert cfg.hard_cuts == int(val)
    if cfg.doing_heavy:
        from tests.test_weight_gpu import sf_table
        orc.set_sf_table(*sf_table())
    sim = Simc(cfg, mode="strict")
    try:
        for arm in sorted({cfg.electron_arm, cfg.hadron_arm}):
            sim.set_optics(load_optics_fixture(arm))
        if cfg.doing_heavy:
            sim.set_sf_table(*sf_table())
        n = 40000
        ref = orc.run(cfg, 0, n, 21, threads=8)
        acc = sim.accum_clear()
        sim.run(0, n, 21, acc)
        accum_equal_exact(acc, ref)
        assert ref.nsuccess > 200, ref.nsuccess
        a, b = acc.wtcontribute.value(), ref.wtcontribute.value()
        assert abs(a - b) <= LOOSE * abs(b), (a, b)
        assert acc.nonfinite == ref.nonfinite == 0
    finally:
        sim.close()
