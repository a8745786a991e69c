"""GPU parity tests of the event loop for D(e,e'p)n (h2.theory) and C(e,e'p) with the independent-particle
spectral function of c12.theory: per-try records, exact accumulators and ntuple rows of libsimc_b200 against
the CPU oracle on the same counter-based random stream; theory tables through both entry points."""
import os

import numpy as np
import pytest

from simc_gfortran_b200 import Simc, config_from_deck, load_optics_fixture
from tests.oracle_lib import load_theory_fixture, write_theory_file
from tests.test_loop_gpu import LOOSE, RECON_LOOSE, RTOL, SCALE, accum_equal_exact, rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECKS = {"h2": "d1_eep_deuterium_hms_sos.inp", "c12": "c2t_eep_carbon_theory_hms_sos.inp"}
SC = SCALE.copy()
SC[5] = 1e-9            # weights: sigma_cc1 x spectral function


@pytest.fixture(scope="module")
def data_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("theory")
    for nm in ("h2", "c12"):
        write_theory_file(load_theory_fixture(nm), str(d / f"{nm}.theory"))
    return str(d)


@pytest.fixture(scope="module", params=["h2", "c12"])
def case(request, oracle_with_optics, data_dir):
    nm = request.param
    cfg = config_from_deck(os.path.join(ROOT, "decks", DECKS[nm]), data_dir=data_dir)[0]
    t = load_theory_fixture(nm)
    oracle_with_optics.set_theory_table(t, cfg.doing_heavy)
    s = Simc(cfg, mode="strict")
    for arm in (1, 2):
        s.set_optics(load_optics_fixture(arm))
    s.set_theory_table(t)
    yield nm, cfg, s, oracle_with_optics, data_dir
    s.close()


def test_event_records(case):
    nm, cfg, sim, orc, _ = case
    n = 40000
    ref, ref_stage = orc.event_batch(cfg, 500, n, 31)
    rec, stage = sim.event_batch(500, n, 31)
    assert np.array_equal(stage, ref_stage), f"{(stage != ref_stage).sum()} tries end at a different stage"
    for k in (0, 2, 3, 4):
        assert np.array_equal(rec[k], ref[k]), sim.event_field_names()[k]
    names = sim.event_field_names()
    gen_ok = stage >= 1
    for k in (13, 14, 17, 18, 26, 27, 28, 29):
        e = rel_err(rec[k][gen_ok], ref[k][gen_ok], SC[k])
        assert e.max() <= RTOL, (names[k], float(e.max()))
    groups = (
        ([7, 8] + list(range(10, 32)) + [35, 36, 37, 47], stage >= 1, LOOSE),
        ([32, 33, 34], stage >= 2, LOOSE),
        ([41, 42, 43], stage >= 2, RECON_LOOSE),
        ([38, 39, 40], stage >= 3, RECON_LOOSE),
        ([1, 5, 6, 9, 44, 45, 46], stage == 4, RECON_LOOSE),
    )
    for fields, mask, tol in groups:
        for k in fields:
            e = rel_err(rec[k][mask], ref[k][mask], SC[k])
            assert e.max() <= tol, (names[k], float(e.max()))
    done = stage == 4
    assert done.sum() > 500 and (stage == 0).sum() > 10
    if nm == "h2":
        assert abs(np.median(rec[44][done]) - 2.2) < 3.0          # Em: the deuteron binding energy
        assert np.all(rec[8][gen_ok] > 0)                          # |dEp'/dEm| x 1/cos^3


def test_accumulators_against_oracle(case):
    nm, cfg, sim, orc, _ = case
    n = 60000
    ref = orc.run(cfg, 0, n, 4, threads=8)
    acc = sim.accum_clear()
    sim.run(0, n, 4, acc)
    accum_equal_exact(acc, ref)
    assert acc.unsupported == ref.unsupported == 0 and acc.nsuccess > 1000
    for f in ("wtcontribute", "sum_sigcc"):
        a, b = getattr(acc, f).value(), getattr(ref, f).value()
        assert abs(a - b) <= RECON_LOOSE * abs(b), f
        assert getattr(acc, f).qexp == getattr(ref, f).qexp
    for k in range(30):
        assert abs(acc.contrib[k].lo - ref.contrib[k].lo) <= 1e-7 * max(1.0, abs(ref.contrib[k].lo)), k
        assert abs(acc.contrib[k].hi - ref.contrib[k].hi) <= 1e-7 * max(1.0, abs(ref.contrib[k].hi)), k


def test_ntuple_rows(case):
    nm, cfg, sim, orc, _ = case
    ref, ref_try = orc.ntuple_batch(cfg, 100, 20000, 12)
    rows, tries = sim.ntuple_batch(100, 20000, 12)
    assert rows.shape == ref.shape and rows.shape[1] == 46 and rows.shape[0] > 300
    assert np.array_equal(tries, ref_try)
    scale = np.ones(46)
    scale[43] = 1e-9
    for k in range(46):
        e = rel_err(rows[:, k], ref[:, k], scale[k])
        assert e.max() <= RECON_LOOSE, (k, float(e.max()))


def test_theory_file_reader_gives_the_same_run(case):
    nm, cfg, sim, orc, data_dir = case
    a = sim.accum_clear()
    sim.run(0, 20000, 2, a)
    s2 = Simc(cfg, mode="strict")
    try:
        for arm in (1, 2):
            s2.set_optics(load_optics_fixture(arm))
        acc = s2.accum_clear()
        with pytest.raises(Exception) as ei:
            s2.run(0, 100, 2, acc)
        assert "theory" in str(ei.value)
        s2.load_theory_file(os.path.join(data_dir, f"{nm}.theory"))
        b = s2.accum_clear()
        s2.run(0, 20000, 2, b)
        assert bytes(a) == bytes(b)
    finally:
        s2.close()
