"""CPU tests that pin the oracle where a pin exists: published known-answer
vectors of the two generators, invariants the reference asserts on its own data files, and
the committed regression vectors."""
import os

import numpy as np
import pytest

from simc_gfortran_b200 import load_optics_fixture
from tests.oracle_lib import transport_inputs

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_ranlux_known_answers(oracle):
    # F. James, CPC 79 (1994) 111: RANLUX default (luxury 3, seed 314159265), numbers 1-5 and 101-105
    u = oracle.ranlux(314159265, 3, 10000)
    np.testing.assert_allclose(u[:5], [0.53981817, 0.76155043, 0.06029940, 0.79600263, 0.30631220], atol=5e-9)
    np.testing.assert_allclose(u[100:105], [0.43156743, 0.03774416, 0.24897110, 0.00147784, 0.90274453], atol=5e-9)
    # GSL's regression value for the same generator (gsl_rng_ranlux, p=223): 10000th output
    assert int(u[9999] * 2 ** 24) == 12077992


def test_philox_known_answers(oracle):
    import ctypes as C
    # Random123 kat_vectors, philox4x32-10
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, want in kat:
        c = np.array(ctr, np.uint32); k = np.array(key, np.uint32); o = np.zeros(4, np.uint32)
        oracle.L.oracle_philox_block(c.ctypes.data_as(C.c_void_p), k.ctypes.data_as(C.c_void_p),
                                     o.ctypes.data_as(C.c_void_p))
        assert list(o) == want


def test_philox_uniform_mapping(oracle):
    u = oracle.philox_uniforms(12345, 7, 1000)
    assert (u > 0).all() and (u < 1).all()
    # draw d comes from words (0,1)/(2,3) of block d/2 at counter (d/2, 0, try_lo, try_hi)
    import ctypes as C
    c = np.array([3, 0, 7, 0], np.uint32); k = np.array([12345, 0], np.uint32); o = np.zeros(4, np.uint32)
    oracle.L.oracle_philox_block(c.ctypes.data_as(C.c_void_p), k.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p))
    k6 = ((int(o[1]) << 32) | int(o[0])) >> 12
    k7 = ((int(o[3]) << 32) | int(o[2])) >> 12
    assert u[6] == (k6 + 0.5) / 2 ** 52 and u[7] == (k7 + 0.5) / 2 ** 52


@pytest.mark.parametrize("arm,n_classes,drifts", [(1, 12, [1, 4, 7, 10, 12]), (5, 32, None), (2, 10, None),
                                                  (3, 12, None), (4, 12, None)])
def test_optics_file_invariants(arm, n_classes, drifts):
    """What the reference checks at load: class counts (hms/mc_hms.f:188, shms/mc_shms.f:344,
    sos/mc_sos.f:158, hrsl/mc_hrsl.f:151) and pure drifts whose length equals the !LENGTH:
    comment (shared/transp.f:449-454)."""
    t = load_optics_fixture(arm)
    assert t.n_classes == n_classes
    if drifts:
        for k in drifts:
            assert t.adrift[k - 1] == 1
            assert abs(t.driftdist[k - 1] - t.length_cm[k - 1]) < 0.01
    for k in range(t.n_classes):
        if t.adrift[k] and t.length_cm[k] > 0:
            assert abs(t.driftdist[k] - t.length_cm[k]) < 0.01
    assert (t.fwd_expon.sum(axis=1) <= 6).all() and (t.fwd_expon >= 0).all()
    # shipped term counts
    if arm == 1:
        n = np.diff(t.class_start)
        assert [int(n[k - 1]) for k in (2, 3, 5, 6, 8, 9, 11)] == [291, 278, 339, 318, 269, 253, 361]
        assert len(t.rec_coeff) == 461


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_fixture_matches_reference_files(oracle):
    for arm, (fwd, rec) in {1: ("hms/forward_cosy.dat", "hms/recon_cosy.dat"),
                            5: ("shms/shms_forward.dat", "shms/shms_recon.dat"),
                            2: ("sos/forward_cosy.dat", "sos/recon_cosy.dat"),
                            3: ("hrsr/hrs_forward_cosy.dat", "hrsr/hrs_recon_cosy.dat"),
                            4: ("hrsl/hrs_forward_cosy.dat", "hrsl/hrs_recon_cosy.dat")}.items():
        oracle.load_optics(arm, "/root/reference/" + fwd, "/root/reference/" + rec)
        a, b = oracle.export_optics(arm), load_optics_fixture(arm)
        for f in ("class_start", "fwd_coeff", "fwd_expon", "length_cm", "adrift", "driftdist", "rec_coeff", "rec_expon"):
            assert np.array_equal(getattr(a, f), getattr(b, f)), f


@pytest.mark.parametrize("arm,name", [(1, "hms"), (5, "shms"), (2, "sos"), (3, "hrsr"), (4, "hrsl")])
def test_oracle_reproduces_golden_vectors(oracle_with_optics, arm, name):
    z = np.load(os.path.join(GOLDEN, f"transport_{name}.npz"))
    out, flags = oracle_with_optics.transport_batch(arm, z["inp"], int(z["seed"]))
    assert np.array_equal(flags, z["flags"])
    assert np.array_equal(out, z["out"])           # same compiler flags, same libm: bit for bit
    assert np.array_equal(z["inp"], transport_inputs(arm, z["inp"].shape[1], int(z["seed"])))


@pytest.mark.parametrize("arm", [1, 5, 2, 3, 4])
def test_forward_then_recon_recovers_the_ray(oracle_with_optics, arm):
    """Physics check that ties forward maps, hut and inverse maps together: with smearing off
    the reconstructed target quantities must equal the thrown ones to optics accuracy."""
    n = 3000
    inp = transport_inputs(arm, n, seed=7)
    inp[0] *= 0.5; inp[4] *= 0.3; inp[5] *= 0.3; inp[1] = 0.0; inp[8] = 0.0
    out, flags = oracle_with_optics.transport_batch(arm, inp, seed=3, ms=False, wcs=False)
    ok = flags == 0
    assert ok.mean() > 0.5
    assert np.abs(out[0][ok] - inp[0][ok]).max() < 0.15           # delta, percent
    assert np.abs(out[1][ok] - inp[4][ok]).max() < 2.5e-3         # xptar
    assert np.abs(out[2][ok] - inp[5][ok]).max() < 2.5e-3         # yptar
    assert np.abs(out[3][ok] - inp[2][ok]).max() < 0.5            # ytar, cm
    # without multiple scattering / smearing only the resmult draw (HMS) consumes random numbers
    assert set(np.unique(out[11][ok])) == ({1.0} if arm in (1, 2) else {0.0})


def test_drift_class_equals_project(oracle_with_optics):
    """transp() on a pure-drift class must reproduce project() (transp.f:399-438): send rays
    through HMS with all apertures passed and compare x at Q1 entrance by hand."""
    t = load_optics_fixture(1)
    k = 0
    sl = slice(t.class_start[k], t.class_start[k + 1])
    co, ex = t.fwd_coeff[sl], t.fwd_expon[sl]
    ray = np.array([0.3, 12.0, -0.4, 7.0, 2.0])       # cm, mrad, cm, mrad, %
    tot = np.zeros(5)
    for c, e in zip(co, ex):
        tot += c * np.prod(ray ** e)
    L = t.driftdist[k]
    assert abs(tot[0] - (ray[0] + ray[1] * 1e-3 * L)) < 1e-9
    assert abs(tot[2] - (ray[2] + ray[3] * 1e-3 * L)) < 1e-9
    assert abs(tot[1] - ray[1]) < 1e-12 and abs(tot[3] - ray[3]) < 1e-12
