"""trg_track.f restated (oracle/field.cpp): the field of the polarised target and the tracking through it.

Known answers that do not come from the oracle itself:
* in trgInit's uniform test field (5 T along the field axis over |z| <= 26 cm, r <= 16 cm) a particle moving
  perpendicular to the field runs on a circle of radius v E / (90 B) = p * 29.9792458 / (90 B) cm -- the constant 90
  is the reference's c^2 in its units (trgXTrack / trgTrackToPlane: factor = 90 / E);
* the measured map (trg_field_map.dat) is an axially symmetric solenoid field: 5.1 T at the centre, Br = 0 on the axis
  and in the mid-plane, div B = 0 and curl B = 0 near the centre to the accuracy of central differences on the 2 cm grid
  (~0.02 T/cm where the gradients themselves are ~0.1 T/cm);
* with no field a track is a straight line, and track_from_tgt ends on the plane z = 100 cm."""
import numpy as np
import pytest

from tests.oracle_lib import load_field_fixture

CC = 29.9792458


def test_uniform_field_circle(oracle):
    oracle.set_field_map(None, None, 0.0, 0.0)            # field axis = z of the spectrometer
    # B = (0, 0, 5) inside the test volume, zero outside
    b = oracle.field_at(1, np.array([[0.0, 3.0, 15.9, 18.5, 0.0, 17.0], [0.0, 4.0, 0.0, 0.0, 0.0, 0.0], [0.0, 10.0, -25.0, 0.0, 28.5, 0.0]]))
    assert np.allclose(b[:, :3], [[0, 0, 0], [0, 0, 0], [5, 5, 5]]) and np.all(b[:, 3:5] == 0)
    assert abs(b[2, 5] - 2.5) < 1e-12                     # bilinear between the last node inside (16 cm) and the first outside
    p, m = 300.0, 0.51099906                              # MeV/c: R = 300 * 29.979 / 450 = 19.99 cm would leave r <= 16 cm ...
    per_step = []
    for p in (100.0, 200.0):                              # ... so stay inside: R = 6.66 cm and 13.3 cm around (R, 0)
        E = np.hypot(p, m)
        v = p / E * CC
        R = v * E / (90.0 * 5.0)
        assert abs(R - p * CC / 450.0) < 1e-12
        # start at the origin moving along +y: positive particle in B = +z turns towards +x (v x B = (vy Bz, 0, 0))
        u0 = np.array([[0.0], [0.0], [0.0], [0.0], [v], [0.0]])
        n = int(2 * np.pi * R)                            # most of a full turn in 1 cm steps ...
        if 2 * R > 15.0:                                  # ... or the arc that stays inside r <= 16 cm (chord 2 R sin(phi/2))
            n = int(R * 2 * np.arcsin(15.0 / (2 * R)))
        traj = oracle.field_steps(1, u0, E, 1.0, n)[:, :, 0]
        r = np.hypot(traj[:, 0] - R, traj[:, 1])
        # classical Runge-Kutta: the radius drifts by ~ (step / R)^5 / 250 cm per step (fourth order globally)
        per_step.append(np.abs(r - R).max() / n)
        assert np.abs(r - R).max() < n * R ** -4 / 500.0, float(np.abs(r - R).max())
        speed = np.sqrt((traj[:, 3:] ** 2).sum(axis=1))
        assert np.abs(speed - v).max() < 1e-5 * v                            # a magnetic field does no work
        assert np.all(traj[:, 2] == 0.0)
        # arc length of 1 cm per step
        phi = np.unwrap(np.arctan2(traj[:, 1], R - traj[:, 0]))
        assert np.allclose(np.diff(phi) * R, 1.0, rtol=1e-4)
        # a negative particle turns the other way
        trn = oracle.field_steps(1, u0, -E, 1.0, n)[:, :, 0]
        assert np.allclose(trn[:, 0], -traj[:, 0], atol=1e-12) and np.allclose(trn[:, 1], traj[:, 1], atol=1e-12)
    assert 24.0 < per_step[0] / per_step[1] < 40.0        # twice the radius: 2^5 less drift per step (local order five)


def test_measured_map_is_a_solenoid(oracle):
    bz, br = load_field_fixture()
    assert bz.shape == (2601,) and abs(bz[0] - 5.1) < 1e-12 and br[0] == 0.0
    oracle.set_field_map(bz, br, 0.0, 0.0)
    BZ, BR = bz.reshape(51, 51), br.reshape(51, 51)       # [ir][iz]
    assert np.all(BR[0, :] == 0.0) and np.all(np.abs(BR[:, 0]) < 1e-6)      # axis, mid-plane
    # the interpolation returns the nodes and is axially symmetric
    zs, rs = np.meshgrid(np.arange(0, 50) * 2.0, np.arange(0, 50) * 2.0)
    pts = np.array([rs.ravel(), np.zeros(rs.size), zs.ravel()])
    b = oracle.field_at(1, pts)
    assert np.allclose(b[2], BZ[:50, :50].ravel(), atol=1e-12) and np.allclose(b[0], BR[:50, :50].ravel(), atol=1e-12)
    rng = np.random.default_rng(1)
    r, z, ph = rng.uniform(0.5, 60, 2000), rng.uniform(-60, 60, 2000), rng.uniform(0, 2 * np.pi, 2000)
    b0 = oracle.field_at(1, np.array([r, 0 * r, z]))
    b1 = oracle.field_at(1, np.array([r * np.cos(ph), r * np.sin(ph), z]))
    assert np.allclose(b1[2], b0[2], rtol=1e-12, atol=1e-13)
    assert np.allclose(np.hypot(b1[0], b1[1]), np.abs(b0[0]), rtol=1e-11, atol=1e-13)
    # Maxwell on the grid: (1/r) d(r Br)/dr + dBz/dz = 0 and dBr/dz - dBz/dr = 0, central differences, inside the coil
    ir, iz = np.meshgrid(np.arange(1, 4), np.arange(1, 8), indexing="ij")      # r <= 6 cm, z <= 14 cm: well inside the coil
    rr = 2.0 * ir
    div = ((rr + 2) * BR[ir + 1, iz] - (rr - 2) * BR[ir - 1, iz]) / (4.0 * rr) + (BZ[ir, iz + 1] - BZ[ir, iz - 1]) / 4.0
    curl = (BR[ir, iz + 1] - BR[ir, iz - 1]) / 4.0 - (BZ[ir + 1, iz] - BZ[ir - 1, iz]) / 4.0
    assert np.abs(div).max() < 0.025 and np.abs(curl).max() < 0.012, (float(np.abs(div).max()), float(np.abs(curl).max()))
    # a rotated field axis (theta = 80 degrees about x): |B| at a point is that of the un-rotated frame
    oracle.set_field_map(bz, br, 80.0, -20.0)
    th = 80.0 * 3.141592653 / 180.0
    pts = rng.uniform(-30, 30, (3, 500))
    be = oracle.field_at(-1, pts)
    x2 = np.sin(th) * pts[2] + np.cos(th) * pts[1]
    x3 = np.cos(th) * pts[2] - np.sin(th) * pts[1]
    oracle.set_field_map(bz, br, 0.0, 0.0)
    bn = oracle.field_at(1, np.array([pts[0], x2, x3]))
    assert np.allclose((be ** 2).sum(axis=0), (bn ** 2).sum(axis=0), rtol=1e-11)


def test_track_from_tgt(oracle):
    bz, br = load_field_fixture()
    rng = np.random.default_rng(5)
    n = 300
    inp = np.array([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), rng.uniform(-2, 2, n), rng.uniform(-0.05, 0.05, n),
                    rng.uniform(-0.03, 0.03, n), rng.choice([-1.0, 1.0], n) * rng.uniform(800, 4000, n),
                    rng.choice([0.51099906, 139.57018, 938.27231], n)])
    # no field: straight lines to z = 100
    oracle.set_field_map(np.zeros(2601), np.zeros(2601), 0.0, 0.0)
    out = oracle.field_batch(1, inp)
    assert np.all(out[5] == 1) and np.allclose(out[2], 100.0, atol=1e-9)
    assert np.allclose(out[0], inp[0] + inp[3] * (100.0 - inp[2]), atol=1e-9) and np.allclose(out[3], inp[3], atol=1e-13)
    assert np.allclose(out[1], inp[1] + inp[4] * (100.0 - inp[2]), atol=1e-9) and np.allclose(out[4], inp[4], atol=1e-13)
    # the measured field at 80 degrees to the spectrometer: the track is bent by ~ 0.3 * int(B dl) / p, opposite for opposite
    # charges, and ends on the plane; the speed is unchanged (checked through the slopes' normalisation upstream)
    oracle.set_field_map(bz, br, 80.0, 80.0)
    pos = inp.copy(); pos[5] = np.abs(pos[5])
    neg = pos.copy(); neg[5] = -pos[5]
    op, on = oracle.field_batch(1, pos), oracle.field_batch(1, neg)
    assert np.all(op[5] == 1) and np.all(on[5] == 1)
    assert np.allclose(op[2], 100.0, atol=1e-9) and np.allclose(on[2], 100.0, atol=1e-9)
    bend_p, bend_n = op[3] - pos[3], on[3] - neg[3]
    assert np.all(bend_p * bend_n < 0) and np.all(np.abs(bend_p) > 1e-3)
    assert np.allclose(bend_p, -bend_n, rtol=0.15)
    # int B dl of a 5 T, ~25 cm coil seen at 80 degrees is ~1 T m: bend ~ 0.3 / p[GeV]
    est = 0.2998 * 1.0 / (pos[5] / 1000.0)
    assert np.all(np.abs(bend_p) < 2.0 * est) and np.all(np.abs(bend_p) > 0.3 * est)
