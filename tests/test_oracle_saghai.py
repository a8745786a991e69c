"""CPU tests of the Saghai kaon model's restatement (physics_kaon.f:241-489 with CERNLIB's fint, cern/fint.f): the
interpolation routine against its defining properties, the library's reader of the reference's table files against
the committed fixture (made by an independent column-wise reading, tools/make_fixtures.py), and the cross section
against a direct evaluation of Saghai's formula from the interpolated amplitudes."""
import os

import numpy as np
import pytest

from tests.oracle_lib import Oracle, load_saghai_fixture

REF = "/root/reference"


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def test_fint_is_multilinear_and_exact_on_nodes(orc):
    rng = np.random.default_rng(4)
    g1 = np.array([0.0, 0.5, 1.25, 2.0, 4.0], np.float32)
    g2 = np.array([-1.0, 0.0, 3.0], np.float32)
    g3 = np.array([10.0, 20.0, 40.0, 80.0], np.float32)
    ent = np.concatenate([g1, g2, g3])
    nent = [5, 3, 4]
    # a function that is linear in each argument is reproduced inside and OUTSIDE the grid (fint extrapolates)
    f = lambda a, b, c: 1.5 + 2.0 * a - 0.25 * b + 0.125 * c + 0.5 * a * b - 0.0625 * b * c + 0.03125 * a * b * c
    tab = np.array([[[f(a, b, c) for a in g1] for b in g2] for c in g3], np.float32)       # a fastest
    for _ in range(200):
        x = [rng.uniform(-0.5, 4.5), rng.uniform(-1.5, 3.5), rng.uniform(5.0, 90.0)]
        x32 = np.array(x, np.float32)
        got = orc.fint(x32, nent, ent, tab)
        assert abs(got - f(*[float(v) for v in x32])) < 2e-5 * max(1.0, abs(got)), (x, got)
    # on a node in every argument: the table value itself, bit for bit
    for i, j, k in ((0, 0, 0), (4, 2, 3), (2, 1, 1), (3, 0, 2)):
        assert orc.fint([g1[i], g2[j], g3[k]], nent, ent, tab) == float(tab[k, j, i])
    # on a node in ONE argument only (the knots of that argument are not split)
    got = orc.fint([g1[2], 1.0, 30.0], nent, ent, tab)
    assert abs(got - f(float(g1[2]), 1.0, 30.0)) < 2e-5 * abs(got)
    # a two-point axis takes the NDIM = 2 branch of the source
    tab2 = np.array([[f(a, b, 0.0) for a in g1] for b in (g2[0], g2[2])], np.float32)
    got = orc.fint([1.0, 1.0], [5, 2], np.concatenate([g1, [g2[0], g2[2]]]), tab2)
    assert abs(got - f(1.0, 1.0, 0.0)) < 2e-5


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is not present on this machine")
@pytest.mark.parametrize("which,name", [(0, "saghai_proton.dat"), (1, "saghai_sigma0.dat")])
def test_library_reader_on_the_reference_files(built_lib, which, name):
    """simc_b200_read_saghai_file on the reference's own file == the fixture, bit for bit."""
    from simc_gfortran_b200.lib import read_saghai_file
    got = read_saghai_file(os.path.join(REF, name), which)
    want = load_saghai_fixture(which)
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    if which == 0:
        # first record of the Lambda file: s = 2.6, Q2 = 0, angle 0 (dbase.f:648-656 order: re, im pairs)
        assert got[0, 0, 0, 0] == np.float32(-0.1155) and got[6, 0, 0, 0] == np.float32(0.0023)
        assert got[3, 0, 0, 0] == np.float32(0.0015) and got[10, 0, 0, 0] == np.float32(0.008)
    else:
        # the Sigma0 file starts with a fragment of a line, so the reference's reads sit one line early: its first
        # "amplitude" read picks up the kinematic line 0.8998E+00 0.7500E+00 0.4531E+00 ... (five numbers + blank)
        assert got[0, 0, 0, 0] == np.float32(0.8998) and got[6, 0, 0, 0] == np.float32(0.75)
        assert got[2, 0, 0, 0] == np.float32(0.5691) and got[8, 0, 0, 0] == 0.0


def test_eekeek_against_saghais_formula(orc):
    """sigma = dsigt + dsigl + dsigp + dsigi from the amplitudes at a grid node (no interpolation involved)."""
    tab = load_saghai_fixture(0)
    orc.set_saghai_table(0, tab)
    Mk, Mp, hbarc, mlam = 493.677, 938.27231, 197.327053, 1115.68
    i_s, i_q, i_a = 3, 4, 6                       # s = 2.6 + 0.3*3 GeV^2, Q2 = 0.8 GeV^2, 60 degrees
    ss = float(np.float32(2.6 + 0.3 + 0.3 + 0.3))
    ss = float(np.float32(np.float64(2.6) + 0.3 + 0.3 + 0.3))
    q22 = float(np.float32(0.0 + 0.2 + 0.2 + 0.2 + 0.2)) * 1e6
    angl = np.deg2rad(60.0)
    theta, phi, eps = 0.05, 0.0, 0.7
    got = orc.saghai_batch(True, mlam, [ss], [q22], [angl], [theta], [phi], [eps])[0]
    z = [complex(float(tab[k, i_a, i_q, i_s]), float(tab[6 + k, i_a, i_q, i_s])) for k in range(6)]
    z1, z2, z3, z4, z7, z8 = z
    w = np.sqrt(ss) * 1000.0
    skc = np.sqrt(max((w * w - Mk ** 2 - mlam ** 2) ** 2 - 4 * Mk ** 2 * mlam ** 2, 0.0)) / 2 / w
    q0 = (q22 + w * w - Mp ** 2) / 2 / Mp
    q0c = (-q22 + q0 * Mp) / w
    qr = np.sqrt(q22) / q0c
    aflx = skc / 2 / w / (w * w - Mp ** 2) * hbarc ** 2 * 1e4
    x, sx = np.cos(angl), np.sin(angl)
    mix = (z1.conjugate() * z4 - z2.conjugate() * z3 + z3.conjugate() * z4 * x).real
    want = (aflx * (abs(z1) ** 2 + abs(z2) ** 2 + 2 * (z1.conjugate() * z2).real * x + 0.5 * sx ** 2 * (abs(z3) ** 2 + abs(z4) ** 2 + 2 * mix))
            + aflx * qr ** 2 * eps * (abs(z7) ** 2 + abs(z8) ** 2 + 2 * (z7.conjugate() * z8).real * x)
            + aflx * eps * np.sin(theta) ** 2 * np.cos(2 * phi) * (0.5 * abs(z3) ** 2 + 0.5 * abs(z4) ** 2 + mix)
            + aflx * np.sqrt(2 * qr ** 2 * eps * (1 + eps)) * np.sin(theta) * np.cos(phi)
            * (z7 * (z3.conjugate() - z2.conjugate() + z4.conjugate() * x) + z8 * (z1.conjugate() + z3.conjugate() * x + z4.conjugate())).real)
    # the float32 angle 60.000004 is not exactly the node 60: allow the interpolation's 1e-7
    assert got > 0 and abs(got - want) < 5e-6 * abs(want), (got, want)
    # smooth and positive over the C5 kinematics (s ~ 3.3 GeV^2, Q2 ~ 0.5 GeV^2, forward angles)
    n = 400
    rng = np.random.default_rng(2)
    out = orc.saghai_batch(True, mlam, rng.uniform(3.0, 3.8, n), rng.uniform(0.3e6, 0.8e6, n), rng.uniform(0.0, 0.6, n),
                           rng.uniform(0.0, 0.1, n), np.zeros(n), rng.uniform(0.5, 0.9, n))
    assert np.all(np.isfinite(out)) and np.all(out > 0) and out.max() < 5.0
