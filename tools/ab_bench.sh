#!/bin/bash
# A/B of two builds of the library on the same box (under gpurun): default vs the build named in $1
#   gpurun -- 'bash tools/ab_bench.sh cytvdn_b200/libcytvdn_b200_r1order.so'
for L in "" "$1" "" "$1"; do
  echo "LIB=${L:-default}"
  CYTVDN_LIB=$L python tools/plain_bench.py --iters 60
  CYTVDN_LIB=$L python tools/iso_bench.py --iters 30 | head -2
done
