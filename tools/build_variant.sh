#!/bin/bash
# build a tuning variant of the library: tools/build_variant.sh <name> <extra nvcc flags...>
set -e
name=$1; shift
cd "$(dirname "$0")/../smeagol_b200/csrc"
out=../../tools/_build/var_$name
mkdir -p $out
for f in api launchers variant_inst hmma_inst scan_inst_0 scan_inst_1 scan_inst_2 scan_inst_3 scan_inst_4 scan_inst_5 scan_inst_6 scan_inst_7; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --fmad=false "$@" -c $f.cu -o $out/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libsmeagol_b200.so $out/*.o
echo built $out/libsmeagol_b200.so
