"""GPU parity of the Saghai kaon model (eekeek / eekeeks with fint, physics_kaon.f:241-489): ntuple column 54
(`saghai` = ntup%sigcm1, results_write.f:168) of H(e,e'K+)Lambda on the HRS pair (C5) and of the Sigma0 channel,
against the oracle with the same tables.  The model never enters the weight, so everything else must be untouched."""
import os
import re

import numpy as np
import pytest

from simc_gfortran_b200 import Simc, config_from_deck, load_optics_fixture
from tests.oracle_lib import load_saghai_fixture

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def deck_with(tmp_path, edits):
    text = open(os.path.join(ROOT, "decks", "c5_eek_hydrogen_hrsl_hrsr.inp")).read()
    for key, val in edits.items():
        pat = re.compile(r"^(\s*" + re.escape(key) + r"\s*=\s*)([^;\n]*)", re.M)
        assert pat.search(text), key
        text = pat.sub(lambda m: m.group(1) + val + "\t", text, count=1)
    path = tmp_path / "kaon.inp"
    path.write_text(text)
    return str(path)


@pytest.mark.parametrize("channel", ["lambda", "sigma0"])
def test_saghai_column(tmp_path, oracle_with_optics, channel):
    orc = oracle_with_optics
    edits = {} if channel == "lambda" else {"which_kaon": "1"}
    cfg = config_from_deck(deck_with(tmp_path, edits))[0]
    which = 0 if cfg.targ.Mrec_struck < 1150.0 else 1
    assert which == (0 if channel == "lambda" else 1)
    tab = load_saghai_fixture(which)
    sim = Simc(cfg, mode="strict")
    try:
        for arm in (4, 3):
            sim.set_optics(load_optics_fixture(arm))
        n = 40000
        rows0, tries0 = sim.ntuple_batch(0, n, 23)                  # no tables: the column is zero
        assert len(rows0) > 50 and np.all(rows0[:, 53] == 0.0)
        sim.set_saghai_table(which, tab)
        orc.set_saghai_table(which, tab)
        rows, tries = sim.ntuple_batch(0, n, 23)
        ref, ref_tries = orc.ntuple_batch(cfg, 0, n, 23)
        assert np.array_equal(tries, ref_tries) and np.array_equal(tries, tries0)
        assert rows.shape[1] == ref.shape[1] == 55
        a, b = rows[:, 53], ref[:, 53]
        assert np.all(np.isfinite(b)) and np.abs(b).max() > 0
        # interpolation weights and the amplitudes are exact products of the same numbers; what differs is the last
        # ulp of sin / cos / sqrt of the vertex quantities that go in (s, Q2, theta_cm are themselves 5e-9 apart
        # between the two sides, tests/test_loop_gpu.py), amplified by the cancellations between the four terms
        scale = np.abs(b).max()
        assert np.abs(a - b).max() <= 2e-7 * scale, float(np.abs(a - b).max() / scale)
        # the weight and every other column do not know about the model
        other = [k for k in range(55) if k != 53]
        assert np.array_equal(rows[:, other], rows0[:, other])
        # and the accumulators of a plain run are those of a run without tables
        acc = sim.run(0, n, 23, sim.accum_clear())
        sim2 = Simc(cfg, mode="strict")
        try:
            for arm in (4, 3):
                sim2.set_optics(load_optics_fixture(arm))
            acc2 = sim2.run(0, n, 23, sim2.accum_clear())
        finally:
            sim2.close()
        assert bytes(acc) == bytes(acc2)
    finally:
        orc.set_saghai_table(which, None)
        sim.close()
