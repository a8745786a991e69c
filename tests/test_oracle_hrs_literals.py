"""SURVEY A.5 made explicit.  The reference's build flags (Makefile:63: -fdefault-real-8 without -fdefault-double-8)
make `...d0` literals REAL(16); hrsl/mc_hrsl.f:248-466 and mc_hrsl_hut.f:293,356 (same in hrsr) pass them by reference
to REAL*8 dummies, which read the low 8 bytes: lengths like -3.1e+232 cm and a chamber tilt of 0 degrees.  The product
(and the oracle by default) implement the literals AS WRITTEN; the oracle's as_built switch reproduces the
reinterpretation so that the difference is on record: which tracks the un-rotated VDC frame accepts, and the path
length (hence the kaon survival probability of C5)."""
import numpy as np

from simc_gfortran_b200 import load_optics_fixture
from tests.oracle_lib import Oracle, transport_inputs


def test_as_built_literals_change_path_length_and_the_vdc_cut():
    orc = Oracle()
    for arm in (3, 4):
        orc.set_optics(load_optics_fixture(arm))
        inp = transport_inputs(arm, 20000, seed=5)
        orc.L.oracle_set_hrs_literals(0)
        a, fa = orc.transport_batch(arm, inp, seed=3)
        orc.L.oracle_set_hrs_literals(1)
        try:
            b, fb = orc.transport_batch(arm, inp, seed=3)
        finally:
            orc.L.oracle_set_hrs_literals(0)
        ok = (fa == 0) & (fb == 0)
        assert ok.sum() > 500
        # as written: a physical path length of ~25 m; as built: the sum of reinterpreted "lengths", ~1e232 cm
        assert np.all((a[8][ok] > 2000.0) & (a[8][ok] < 3000.0))
        assert np.all(np.abs(b[8][ok]) > 1e200)
        # the magnet maps do not read zd, so the track itself is the same up to the hut ...
        assert np.array_equal(a[0][ok], b[0][ok]) or np.abs(a[0][ok] - b[0][ok]).max() < 1e-9
        # ... where the un-rotated VDC frame (45.0d0 -> 0 degrees) accepts a different set of tracks
        assert (fa != fb).sum() > 0
