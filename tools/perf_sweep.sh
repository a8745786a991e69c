#!/bin/bash
# Builds variants of libsimc_b200.so with different compile-time knobs into build_variants/ (HERE, nvcc
# cross-compiles) so that one gpurun call can bench them all:
#   tools/perf_sweep.sh build "name1:-DKNOB=1 -DOTHER=2" "name2:..."
#   gpurun -- 'tools/perf_sweep.sh run name1 name2'
set -e
cd "$(dirname "$0")/.."
ROOT=$PWD
mode=$1; shift
if [ "$mode" = build ]; then
  mkdir -p build_variants
  for spec in "$@"; do
    name=${spec%%:*}; flags=${spec#*:}
    d=build_variants/obj_$name
    rm -rf $d; mkdir -p $d
    cp simc_gfortran_b200/csrc/*.cu simc_gfortran_b200/csrc/*.cuh simc_gfortran_b200/csrc/*.h simc_gfortran_b200/csrc/*.cpp simc_gfortran_b200/csrc/Makefile $d/
    # the Makefile's include path is relative to csrc: ../../include
    sed -i "s|-I../../include|-I$ROOT/include|; s|../../include/simc_b200.h|$ROOT/include/simc_b200.h|" $d/Makefile
    sed -i "s|\"../../include/simc_b200.h\"|\"$ROOT/include/simc_b200.h\"|" $d/*.cu $d/*.cuh $d/*.h $d/*.cpp
    ( make -C $d XFLAGS="$flags" OUT=$ROOT/build_variants/lib_$name.so > $d/build.log 2>&1 && echo "built $name" && grep -h "Function properties for _ZN4simc6strict10k_generate\|k_armILi1ELi1\|Used" $d/ptxas_strict.log | grep -A1 "k_generate\|k_arm" | grep Used | head -3 ) &
  done
  wait
  rm -rf build_variants/obj_*
else
  mkdir -p gpurun_out
  for name in "$@"; do
    lib=$ROOT/build_variants/lib_$name.so
    [ "$name" = base ] && lib=$ROOT/simc_gfortran_b200/libsimc_b200.so
    echo "== $name"
    SIMC_B200_LIB=$lib python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python tools/bench_line.py $name | tee -a gpurun_out/sweep.log
  done
fi
