"""Known answers the reference itself holds as DATA (not code) for the Hall A spectrometers, checked on the oracle:

* `drifts.txt` (12 rows: element length, running total; the same numbers as the `!LENGTH:` comments of
  `hrsl/hrs_forward_cosy.dat` and `hrsr/hrs_forward_cosy.dat` and as the "exit is z=..." comments of
  `hrsl/mc_hrsl.f:270,329,411,477`): the central ray's path length through HRS-L / HRS-R must be the table's total plus the
  hut's last plane (`hcal_4ta_zpos = 407.3 + 25.0`, `hrsl/mc_hrsl_hut.f:181`).  This also bears on SURVEY A.5 (the
  `62.75333333d0`-style literals): the table carries the values AS WRITTEN, which is what product and oracle use.
* `hrsr/hrs_aperture_info.txt`: the two circular apertures in front of the slit box (r = 7.3787 cm at 65.686 cm and
  r = 7.4092 cm at 80.436 cm from the pivot; `hrsl/mc_hrsl.f:159-179`) lie in the field-free region, so straight rays
  on either side of the radius are a known answer for the accept/reject flag.

The GPU twin (the product through the C ABI) is tests/test_zz_reference_tables_gpu.py.
"""
import os

import numpy as np
import pytest

from tests.oracle_lib import ARM_KIN

# drifts.txt:1-12 (element length, running total; cm)
HRS_DRIFTS = [(159.03000, 159.03000), (62.75333, 221.78333), (31.37667, 253.16000), (117.20000, 370.36000),
              (121.77333, 492.13333), (60.88667, 553.02000), (443.08000, 996.10000), (659.73446, 1655.83446),
              (159.25000, 1815.08446), (121.78667, 1936.87112), (60.89333, 1997.76446), (345.23000, 2342.99446)]
HRS_HUT_LAST_PLANE = 407.3 + 25.0          # hrsl/mc_hrsl_hut.f:181, hrsr/mc_hrsr_hut.f: hcal_4ta_zpos
# hrsr/hrs_aperture_info.txt: (distance from the pivot, radius), cm
HRS_FRONT_CIRCLES = [(65.686, 7.3787), (80.436, 7.4092)]
REF = "/root/reference"


def central_ray(arm, n=1):
    inp = np.zeros((9, n))
    inp[6] = ARM_KIN[arm][1]
    inp[7] = ARM_KIN[arm][0]
    return inp


def test_the_table_is_the_reference_file():
    """The constants above are what drifts.txt holds (checked where the reference is present: this container)."""
    run = 0.0
    for length, total in HRS_DRIFTS:
        run += length
        assert abs(run - total) < 2e-5
    path = os.path.join(REF, "drifts.txt")
    if not os.path.exists(path):
        pytest.skip("reference not present (GPU box)")
    rows = [tuple(float(x) for x in line.split()) for line in open(path) if line.strip()]
    assert rows == HRS_DRIFTS
    # the optics file's own canonical lengths (metres) are the same twelve numbers
    for arm_dir in ("hrsl", "hrsr"):
        lengths = [float(line.split()[1]) for line in open(os.path.join(REF, arm_dir, "hrs_forward_cosy.dat"))
                   if line.startswith("!LENGTH")]
        assert len(lengths) == 12
        for got, (want, _) in zip(lengths, HRS_DRIFTS):
            assert abs(100.0 * got - want) < 1e-5


@pytest.mark.parametrize("arm", [3, 4])
def test_central_ray_path_length(oracle_with_optics, arm):
    out, flags = oracle_with_optics.transport_batch(arm, central_ray(arm), seed=1, ms=False, wcs=False)
    assert flags[0] == 0
    assert abs(out[8, 0] - (HRS_DRIFTS[-1][1] + HRS_HUT_LAST_PLANE)) < 1e-5


@pytest.mark.parametrize("arm", [3, 4])
def test_front_apertures(oracle_with_optics, arm):
    """Vertical rays (x is the dispersive coordinate: the slit's own vertical check, a different STOP counter, is the next
    thing an accepted ray meets) just inside and just outside each circle."""
    eps = 1e-6
    inp = central_ray(arm, 4)
    (z1, r1), (z2, r2) = HRS_FRONT_CIRCLES
    assert r1 < r2 and r2 / z2 < r1 / z1      # a parallel ray tests the first circle alone, a ray from the origin the second
    for k, side in enumerate((1 - eps, 1 + eps)):
        inp[1, k] = r1 * side                 # x = r1 * side all the way (dxdz = 0)
        inp[4, 2 + k] = r2 * side / z2        # from the origin: x = r2 * side at z2, 0.82 r1 at z1
    out, flags = oracle_with_optics.transport_batch(arm, inp, seed=1, ms=False, wcs=False)
    in1, out1, in2, out2 = (int(f) for f in flags)
    assert out1 != 0 and out2 != 0
    assert out1 == out2                       # both circles count as lSTOP_slit_hor (mc_hrsl.f:163,174)
    assert in1 != out1 and in2 != out2        # a ray inside the circle gets further (and stops on the slit's vertical edge)
    assert in1 == in2 and in1 != 0
