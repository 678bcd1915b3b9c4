#!/bin/bash
# Development: build a variant of the library with extra flags into build/<name>/libfemhelmholtz_b200.so (load it with
# FH_LIB=build/<name>/libfemhelmholtz_b200.so).   tools/build_variant.sh stat -DFH_PROBE_STAT
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/build/$name
mkdir -p $out
cd $root/fem_helmholtz_eqn_b200/csrc
for f in *.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden "$@" -c $f -o $out/${f%.cu}.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libfemhelmholtz_b200.so $out/*.o -lcudart
echo $out/libfemhelmholtz_b200.so
