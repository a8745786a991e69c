"""CPU tests of the boundary: the library builds, loads, exports every symbol the header
declares, and refuses to compute without a GPU (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from simc_gfortran_b200 import lib as simlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "simc_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(simc_b200_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(built_lib):
    L = C.CDLL(built_lib)
    names = declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_struct_layouts_match(built_lib):
    L = simlib.load_library()          # raises on a sizeof mismatch
    assert L.simc_b200_abi_version() == simlib.ABI_VERSION
    assert L.simc_b200_sizeof(0) == C.sizeof(simlib.RunConfig)
    assert L.simc_b200_sizeof(1) == C.sizeof(simlib.Accum)


def test_stop_names(built_lib):
    L = simlib.load_library()
    assert L.simc_b200_stop_name(1, 0) == b"ok"
    assert L.simc_b200_stop_name(1, 17) == b"scin"
    assert L.simc_b200_stop_name(5, 1) == b"HB_in"
    assert L.simc_b200_stop_name(5, 42) == b"cal_fid"


def test_no_cpu_fallback(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(simlib.SimcError) as ei:
        simlib.Simc()
    assert ei.value.code == -3 and "no CPU fallback" in str(ei.value)


def test_oracle_is_not_linked_into_the_product(built_lib):
    """The product may not route through oracle/: no oracle symbol, no liboracle dependency."""
    import subprocess
    nm = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True).stdout
    assert "oracle_" not in nm and "simc_oracle" not in nm
    ldd = subprocess.run(["ldd", built_lib], capture_output=True, text=True).stdout
    assert "liboracle" not in ldd
    for root, _, files in os.walk(os.path.join(ROOT, "simc_gfortran_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "oracle/" not in src.replace("oracle/: ", "") and "oracle_lib" not in src, f
