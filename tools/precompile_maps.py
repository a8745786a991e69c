#!/usr/bin/env python
"""Runs the map compiler (csrc/mapgen.h + NVRTC; no GPU needed) on the shipped optics and leaves the cubins in the
cache next to libsimc_b200.so, so that a GPU box loads them without compiling:
    tools/precompile_maps.py [cache_dir] [dump_dir]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simc_gfortran_b200 import load_optics_fixture, precompile_optics


def main(cache=None, dump=None, quiet=False):
    for arm in (1, 5, 2, 3, 4):
        t = load_optics_fixture(arm)
        for strict in (True, False):
            t0 = time.time()
            name = "strict" if strict else "fast"
            info = precompile_optics(t, strict, cache, os.path.join(dump, f"maps_arm{arm}_{name}.cu") if dump else None)
            if not quiet:
                print(f"arm {arm} {name}: {info[0]} stretches, source {info[1]} B, cubin {info[2]} B, "
                      f"{'cached' if info[3] else 'compiled'} in {time.time() - t0:.1f} s")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None, sys.argv[2] if len(sys.argv) > 2 else None)
