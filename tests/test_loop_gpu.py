"""GPU parity tests of the whole event loop (simc.f:169-351) for C1 = H(e,e'p), HMS + SHMS:
per-try records and the exact integer accumulators of libsimc_b200 against the CPU oracle on
the same counter-based random stream."""
import ctypes as C
import os

import numpy as np
import pytest

from simc_gfortran_b200 import Accum, Simc, config_from_deck, load_optics_fixture

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECK = os.path.join(ROOT, "decks", "c1_eep_hydrogen_hms_shms.inp")
RTOL = 1e-12
# The reference's soft-photon formulas (bremos/inter, brem.f:216-240,411-418) form ar1 ~ 1e-8 as
# 0.5 - 0.4999..., so their OUTPUT moves by ~1e-10 relative when an INPUT moves by one ulp
# (tests/test_oracle_event.py::test_reference_radiative_weight_is_ill_conditioned measures this on
# the oracle itself).  A GPU libm that differs from glibc in the last bit of one sin/cos/acos
# therefore shifts everything downstream of the radiative constants at that level -- and, through
# the REAL*4 drift-chamber coordinates, recon quantities of a few % of events by one float ulp.
# Whole-event records are hence compared at LOOSE; the 1e-12 bar is enforced stage by stage on
# identical inputs (test_radc_stage, tests/test_transport_gpu.py).
LOOSE = 5e-9
RECON_LOOSE = 2e-5

# record fields (simc_b200_event_field_name) that are defined once a try reached a stage
F_ALWAYS = [0, 2, 3, 4]
F_GEN = [7, 8] + list(range(10, 32)) + [35, 36, 37, 47]
F_PARM = [41, 42, 43]
F_EARM_ENTERED = [32, 33, 34]
F_EARM = [38, 39, 40]
F_DONE = [1, 5, 6, 9, 44, 45, 46]
# natural scales for columns that pass through zero
SCALE = np.ones(60)
for k in (13, 14, 17, 18, 33, 34, 36, 37, 39, 40, 42, 43):
    SCALE[k] = 1e-2           # angles
for k in (5, 6, 9):
    SCALE[k] = 1e-6           # cross sections (ub/sr), weights
for k in (44, 45):
    SCALE[k] = 10.0           # Em, Pm (MeV): differences of ~GeV quantities
SCALE[22:25] = 1.0


@pytest.fixture(scope="module")
def cfg():
    return config_from_deck(DECK)[0]


@pytest.fixture(scope="module")
def orc(oracle_with_optics):
    return oracle_with_optics


@pytest.fixture(scope="module", params=["strict", "fast"])
def sim(request, cfg):
    s = Simc(cfg, mode=request.param)
    for arm in (1, 5):
        s.set_optics(load_optics_fixture(arm))
    yield s
    s.close()


def rel_err(a, b, scale):
    return np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), scale)


def test_event_records(sim, orc, cfg):
    n = 20000
    ref, ref_stage = orc.event_batch(cfg, 1000, n, 2024)
    rec, stage = sim.event_batch(1000, n, 2024)
    assert np.array_equal(stage, ref_stage), f"{(stage != ref_stage).sum()} tries end at a different stage"
    for k in F_ALWAYS:
        assert np.array_equal(rec[k], ref[k]), sim.event_field_names()[k]
    names = sim.event_field_names()
    # quantities that do not depend on the radiative constants: tight in the strict variant (the
    # fast variant contracts multiply-adds everywhere, which the ill-conditioned steps amplify)
    tight = RTOL if sim.mode == "strict" else LOOSE
    for k in (8, 13, 14, 26, 27, 28, 29):
        e = rel_err(rec[k][stage >= 1], ref[k][stage >= 1], SCALE[k])
        assert e.max() <= tight, (names[k], float(e.max()))
    no_tail1 = (stage >= 1) & (ref[25] != 1)              # vertex untouched by radiation
    for k in (10, 11, 12, 15, 16, 19, 30, 31):
        e = rel_err(rec[k][no_tail1], ref[k][no_tail1], SCALE[k])
        assert e.max() <= tight, (names[k], float(e.max()))
    # the proton's spectrometer angles come from sqrt(1/cos^2 - 1 - dx^2) (event.f:1640): a
    # cancellation that turns the last-ulp difference of one cos() into ~1e-13 rad
    for k in (17, 18):
        e = np.abs(rec[k][no_tail1] - ref[k][no_tail1])
        assert e.max() <= 1e-11 * (tight / RTOL), (names[k], float(e.max()))
    for fields, mask, tol in ((F_GEN, stage >= 1, LOOSE), (F_EARM_ENTERED, stage >= 2, LOOSE),
                              (F_PARM, stage >= 2, RECON_LOOSE), (F_EARM, stage >= 3, RECON_LOOSE),
                              (F_DONE, stage == 4, RECON_LOOSE)):
        for k in fields:
            e = rel_err(rec[k][mask], ref[k][mask], SCALE[k])
            assert e.max() <= tol, (names[k], float(e.max()))
            if tol == RECON_LOOSE:      # float-ulp flips must stay rare
                assert (e > LOOSE).mean() < 0.1, (names[k], float((e > LOOSE).mean()))
    assert (stage == 4).sum() > 2000 and (stage == 0).sum() > 100 and (stage == 1).sum() > 5000


def test_radc_stage(sim, orc, cfg):
    """radc_init_ev + peaked_rad_weight + sigep on identical dumped vertex vectors: 1e-12."""
    n = 50000
    rec, stage = orc.event_batch(cfg, 0, n, 5)
    m = stage >= 1
    Ein, eE = rec[10][m], rec[11][m]
    rng = np.random.default_rng(3)
    k = m.sum()
    # rebuild the vertex vectors the radiative routines read (elastic kinematics, event.f:518-562)
    Mp = 938.27231
    cth = 1.0 - (Ein / eE - 1.0) * Mp / Ein
    eth = np.arccos(cth)
    phi = rng.uniform(4.6, 4.8, k)
    ue = np.stack([np.sin(eth) * np.cos(phi), np.sin(eth) * np.sin(phi), np.cos(eth)])
    nu = Ein - eE
    q = np.sqrt(2 * Ein * eE * (1 - ue[2]) + nu * nu)
    up = np.stack([-eE * ue[0] / q, -eE * ue[1] / q, (Ein - eE * ue[2]) / q])
    pE = np.sqrt(q * q + Mp * Mp)
    emax = rng.uniform(20.0, 1200.0, k)
    emin = np.where(rng.uniform(size=k) < 0.5, rng.uniform(-50.0, 0.0, k), rng.uniform(0.0, 0.9, k) * emax)
    inp = np.stack([Ein, eE, eth, ue[0], ue[1], ue[2], pE, q, up[0], up[1], up[2], rng.uniform(2e-3, 3e-2, k),
                    rng.uniform(5e-3, 5e-2, k), rng.uniform(0.0, 1.0, k) * emax, emin, emax])
    ref = orc.radc_batch(cfg, inp)
    out = sim.radc_batch(inp)
    err = np.abs(out - ref) / np.maximum(np.abs(ref), 1e-300)
    # columns 23-25 are the on-shell `brem` (brem.f:6-214), which this deck does not use: it rebuilds the scattering
    # angle as 2*asin(sqrt(..)) and feeds it to the same ill-conditioned interference terms, so a last-ulp difference
    # between the CUDA and glibc asin/cos (a few % of the arguments) shows at 1e-9; tests/test_rad_options_gpu.py
    assert err[23:].max() < 1e-7
    err = err[:23]
    # The reference writes the energies as (...)**0.5 (brem.f:383-384,398): that is glibc pow(x,0.5),
    # which differs from the correctly rounded sqrt in the last bit for ~2e-4 of the arguments, and
    # bremos amplifies one ulp of k_f%e by ~1e7.  Those rows are allowed, and counted.
    tol = RTOL if sim.mode == "strict" else 1e-10
    outliers = (err > tol).any(axis=0)
    assert outliers.mean() < 1e-3, float(outliers.mean())
    assert err[:, ~outliers].max() <= tol
    assert err.max() < 1e-7


def accum_equal_exact(a: Accum, b: Accum):
    for f in ("ntried", "nsuccess", "ncontribute", "npasscuts", "ncontribute_no_rad_proton"):
        assert getattr(a, f) == getattr(b, f), f
    assert np.array_equal(np.ctypeslib.as_array(a.hist_n), np.ctypeslib.as_array(b.hist_n))
    assert np.array_equal(np.ctypeslib.as_array(a.stop), np.ctypeslib.as_array(b.stop))
    assert np.array_equal(np.ctypeslib.as_array(a.transp_calls), np.ctypeslib.as_array(b.transp_calls))


def test_accumulators_against_oracle(sim, orc, cfg):
    n = 30000
    ref = orc.run(cfg, 0, n, 777, threads=8)
    acc = sim.accum_clear()
    sim.run(0, n, 777, acc)
    accum_equal_exact(acc, ref)            # counters, count histograms, STOP counters: exact
    assert acc.nsuccess > 3000
    # weight sums: every event weight agrees to <=1e-12, so do the exact fixed-point sums
    for name in ("wtcontribute", "sum_sigcc"):
        a, b = getattr(acc, name).value(), getattr(ref, name).value()
        assert abs(a - b) <= LOOSE * abs(b), name
        assert getattr(acc, name).qexp == getattr(ref, name).qexp
    hw = np.array([[acc.hist_w[k][b].value() for b in range(50)] for k in range(6)])
    hr = np.array([[ref.hist_w[k][b].value() for b in range(50)] for k in range(6)])
    # a float-ulp flip can move an event weight to the neighbouring bin: compare the totals tightly,
    # the bins loosely
    assert np.allclose(hw.sum(axis=1), hr.sum(axis=1), rtol=LOOSE, atol=0)
    assert np.allclose(hw, hr, rtol=1e-3, atol=2e-5 * hr.max())
    for k in range(8):
        assert abs(acc.sumerr[k].value() - ref.sumerr[k].value()) < 1e-4
        assert abs(acc.sumerr2[k].value() - ref.sumerr2[k].value()) < 1e-4
    for k in range(30):
        assert abs(acc.contrib[k].lo - ref.contrib[k].lo) <= 1e-7 * max(1.0, abs(ref.contrib[k].lo)), k
        assert abs(acc.contrib[k].hi - ref.contrib[k].hi) <= 1e-7 * max(1.0, abs(ref.contrib[k].hi)), k
    for k in range(8):
        assert abs(acc.slop[k].lo - ref.slop[k].lo) < 1e-5 and abs(acc.slop[k].hi - ref.slop[k].hi) < 1e-5


def test_run_is_independent_of_batching_and_order(sim):
    """Integer accumulators + counter-based stream: any split of the try range gives the same bits."""
    n = 50000
    a = sim.accum_clear()
    sim.set_batch(1 << 20)
    sim.run(5, n, 99, a)
    b = sim.accum_clear()
    sim.set_batch(4096)
    sim.run(5, 20000, 99, b)
    sim.run(20005, n - 20000, 99, b)
    sim.set_batch(1 << 20)
    assert bytes(a) == bytes(b)


def test_two_handles_with_different_seeds_interleaved(cfg):
    """The generator's round keys are device-wide constant memory (philox.cuh): two handles on one device that run with
    different seeds, their launches interleaved (asynchronous runs in small batches, also from two threads), must each
    see their own stream -- a change of seed waits for the work in flight."""
    import threading
    n = 40000
    sims = []
    try:
        for _ in range(2):
            s = Simc(cfg, mode="strict")
            for arm in (1, 5):
                s.set_optics(load_optics_fixture(arm))
            s.set_batch(4096)
            sims.append(s)
        want = []
        for s, seed in zip(sims, (11, 12)):
            a = s.accum_clear()
            s.run(0, n, seed, a)
            want.append(bytes(a))
        assert want[0] != want[1]
        # one thread, launches alternating between the handles
        accs = [s.accum_clear() for s in sims]            # (also sets the host copy's ranges to empty)
        for k in range(0, n, 8000):
            sims[0].run_async(k, 8000, 11)
            sims[1].run_async(k, 8000, 12)
        got = [bytes(s.fetch(a)) for s, a in zip(sims, accs)]
        assert got == want
        # two threads, one handle each
        out = [None, None]

        def work(i, seed):
            a = sims[i].accum_clear()
            sims[i].run(0, n, seed, a)
            out[i] = bytes(a)

        th = [threading.Thread(target=work, args=(i, seed)) for i, seed in enumerate((11, 12))]
        for t in th:
            t.start()
        for t in th:
            t.join()
        assert out == want
    finally:
        for s in sims:
            s.close()


def test_empty_and_errors(cfg):
    from simc_gfortran_b200 import SimcError
    s = Simc(cfg)
    try:
        with pytest.raises(SimcError) as e:       # optics not loaded
            s.run(0, 10, 1, s.accum_clear())
        assert e.value.code == -4
        for arm in (1, 5):
            s.set_optics(load_optics_fixture(arm))
        acc = s.accum_clear()
        s.run(0, 0, 1, acc)
        assert acc.ntried == 0 and acc.contrib[0].lo == 1e10
    finally:
        s.close()
    import copy
    bad = copy.copy(cfg)
    bad = type(cfg).from_buffer_copy(bytes(cfg))
    bad.doing_hyd_elast = 0
    bad.doing_pion = 1
    s = Simc(bad)
    try:
        with pytest.raises(SimcError) as e:
            s.run(0, 10, 1, s.accum_clear())
        assert "H(e,e'p)" in str(e.value)
    finally:
        s.close()


def test_full_size_properties(sim, cfg):
    """C1 at its BASELINE size (1e6 tries): normalised yield and acceptance are stable between
    two independent streams within statistics; histogram totals are consistent with counters."""
    accs = []
    for seed in (1, 2):
        a = sim.accum_clear()
        sim.run(0, 1000000, seed, a)
        accs.append(a)
    for a in accs:
        assert a.ntried == 1000000
        assert 0.18 < a.nsuccess / a.ntried < 0.24
        geni = np.ctypeslib.as_array(a.hist_n)[2]
        assert geni[0].sum() <= a.ntried and geni[1].sum() == a.ntried      # yptar is generated inside its axis
        gen = np.ctypeslib.as_array(a.hist_n)[1]
        assert gen[1].sum() == a.nsuccess
        assert a.stop[1][0] >= a.stop[0][0] == a.stop[1][1]                # E arm is entered iff the P arm succeeded
        assert a.stop[0][1] == a.nsuccess
    y = [a.wtcontribute.value() / a.ntried for a in accs]
    assert abs(y[0] - y[1]) / y[0] < 0.02
