"""The device logarithm (simc_gfortran_b200/csrc/fastlog.cuh) evaluated on the host (the same header, fma from the
CPU) against libquadmath on 4e6 arguments: (0,1) like the Gaussians of gauss1.f, around 1 where the result goes
through zero, 17 decades, and random positive normals.  Must stay within 0.52 ulp (glibc's own log, which the
reference build calls, measures 0.518 in the same run); special arguments go to the library."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fastlog_accuracy_against_quad_precision(tmp_path):
    exe = str(tmp_path / "fastlog_check")
    subprocess.run(["g++", "-O2", "-mfma", "-ffp-contract=off", "-I", os.path.join(ROOT, "simc_gfortran_b200", "csrc"), "-x", "c++",
                    os.path.join(ROOT, "tests", "aux", "fastlog_host_check.cpp"), "-o", exe, "-lquadmath"], check=True)
    out = subprocess.run([exe, "1000000"], check=True, capture_output=True, text=True).stdout
    m = re.search(r"max ulp: log ([\d.]+) log10 ([\d.]+) \(glibc log ([\d.]+)\)", out)
    assert m, out
    log_ulp, log10_ulp, glibc_ulp = (float(x) for x in m.groups())
    assert log_ulp < 0.52 and log10_ulp < 0.52, out
    assert glibc_ulp < 0.53
    assert "special: -inf -nan inf nan -744.44" in out or "special: -inf nan inf nan -744.44" in out, out


def test_table_is_what_the_generator_writes(tmp_path):
    """The committed table equals a fresh run of tools/gen_fastlog_table.py (mpmath, 200 bits)."""
    import importlib.util
    hdr = os.path.join(ROOT, "simc_gfortran_b200", "csrc", "fastlog_table.h")
    fresh = str(tmp_path / "fastlog_table.h")            # (never rewrite the tracked header: its mtime drives the CUDA build)
    os.environ["SIMC_FASTLOG_OUT"] = fresh
    try:
        spec = importlib.util.spec_from_file_location("gen_fastlog_table", os.path.join(ROOT, "tools", "gen_fastlog_table.py"))
        spec.loader.exec_module(importlib.util.module_from_spec(spec))
    finally:
        del os.environ["SIMC_FASTLOG_OUT"]
    assert open(fresh).read() == open(hdr).read()
