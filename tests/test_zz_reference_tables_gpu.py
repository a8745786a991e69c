"""GPU twin of tests/test_reference_tables_cpu.py: the product (libsimc_b200 through the C ABI) on the known answers the
reference holds as data for the Hall A spectrometers -- the central ray's path length against `drifts.txt` (+ the hut's
last plane, hrsl/mc_hrsl_hut.f:181) and the two circular apertures in front of the slit box against
`hrsr/hrs_aperture_info.txt` (hrsl/mc_hrsl.f:159-179).  The flags must also be the oracle's."""
import numpy as np
import pytest

from simc_gfortran_b200 import Simc, load_optics_fixture
from tests.test_reference_tables_cpu import HRS_DRIFTS, HRS_FRONT_CIRCLES, HRS_HUT_LAST_PLANE, central_ray

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("arm", [3, 4])
def test_hrs_known_answers(oracle_with_optics, arm):
    eps = 1e-6
    inp = central_ray(arm, 5)                 # row 0: the central ray
    (z1, r1), (z2, r2) = HRS_FRONT_CIRCLES
    for k, side in enumerate((1 - eps, 1 + eps)):
        inp[1, 1 + k] = r1 * side             # parallel rays at x = r1 * side: the first circle alone
        inp[4, 3 + k] = r2 * side / z2        # rays from the origin with x = r2 * side at the second circle
    s = Simc(mode="strict")
    try:
        s.set_optics(load_optics_fixture(arm))
        out, flags = s.transport_batch(arm, inp, 1, ms_flag=False, wcs_flag=False)
    finally:
        s.close()
    ref_out, ref_flags = oracle_with_optics.transport_batch(arm, inp, seed=1, ms=False, wcs=False)
    assert np.array_equal(flags, ref_flags)
    assert flags[0] == 0
    assert abs(out[8, 0] - (HRS_DRIFTS[-1][1] + HRS_HUT_LAST_PLANE)) < 1e-5
    in1, out1, in2, out2 = (int(f) for f in flags[1:])
    assert out1 != 0 and out1 == out2 and in1 != out1 and in2 != out2
