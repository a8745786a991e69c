"""The device logarithm on the GPU (simc_b200_log_batch): bit-identical to the same header evaluated on the host
(both sides are IEEE fma arithmetic on the same table), and within 0.52 ulp of mpmath on a sample."""
import os
import subprocess

import numpy as np
import pytest

from simc_gfortran_b200 import Simc, config_from_deck

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HOST = r"""
#include <cstdio>
#include <cstring>
#include <cmath>
#include "fastlog.cuh"
int main(int argc, char** argv) {
  FILE* f = fopen(argv[1], "rb"); FILE* g = fopen(argv[2], "wb");
  double x;
  while (fread(&x, 8, 1, f) == 1) { double y[2] = {simc::fastlog::log(x), simc::fastlog::log10(x)}; fwrite(y, 8, 2, g); }
  fclose(f); fclose(g); return 0;
}
"""


def test_device_log_equals_host_evaluation_and_mpmath(tmp_path):
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(0.0, 1.0, 400000), 0.99 + 0.02 * rng.uniform(size=200000),
                        np.exp(40.0 * (rng.uniform(size=200000) - 0.5)),
                        np.array([1.0, 0.6875, 1.375, np.nextafter(1.0, 0.0), np.nextafter(1.0, 2.0), 2.0, 0.5, 5e-324, 1e-310])])
    src = tmp_path / "h.cpp"
    src.write_text(HOST)
    exe = str(tmp_path / "h")
    subprocess.run(["g++", "-O2", "-mfma", "-ffp-contract=off", "-I", os.path.join(ROOT, "simc_gfortran_b200", "csrc"), str(src), "-o", exe],
                   check=True)
    x.tofile(str(tmp_path / "x.bin"))
    subprocess.run([exe, str(tmp_path / "x.bin"), str(tmp_path / "y.bin")], check=True)
    host = np.fromfile(str(tmp_path / "y.bin")).reshape(-1, 2)
    sim = Simc(config_from_deck(os.path.join(ROOT, "decks", "c1_eep_hydrogen_hms_shms.inp"))[0], mode="strict")
    try:
        ln, l10 = sim.log_batch(x)
        z, _ = sim.log_batch(np.array([0.0, -1.0, np.inf]))
    finally:
        sim.close()
    assert np.array_equal(ln, host[:, 0])
    assert np.array_equal(l10, host[:, 1])
    assert z[0] == -np.inf and np.isnan(z[1]) and z[2] == np.inf
    import mpmath as mp
    mp.mp.prec = 120
    worst = 0.0
    for k in rng.integers(0, x.size - 2, 3000):
        want = mp.log(mp.mpf(float(x[k])))
        if want == 0:
            assert ln[k] == 0.0
            continue
        ulp = mp.mpf(2) ** (mp.floor(mp.log(abs(want), 2)) - 52)
        worst = max(worst, float(abs(mp.mpf(float(ln[k])) - want) / ulp))
    assert worst < 0.52, worst
