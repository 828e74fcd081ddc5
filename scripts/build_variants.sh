#!/bin/bash
# Kernel-variant copies of libgymcook.so for A/B timing on the GPU box (GC_LIBGYMCOOK=<path> selects one).
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
build() {  # name, extra nvcc flags
  name=$1; shift
  rm -f build/variants/libgymcook_$name.so
  GC_LIBGYMCOOK=$PWD/build/variants/libgymcook_$name.so GC_NVCC_EXTRA="$*" python gym-cooking_b200/build.py > /dev/null
  echo built $name: "$@"
}
build c7 -DGC_STEP2_MIN_CTAS_ALL=7 &

build t128 -DGC_STEP2_THREADS=128 -DGC_STEP2_MIN_CTAS_ALL=12 &


wait
build l2a -DGC_STEP2_L2_AHEAD=1 &
build c7l2a -DGC_STEP2_L2_AHEAD=1 -DGC_STEP2_MIN_CTAS_ALL=7 &
wait
build c8 -DGC_STEP2_MIN_CTAS_ALL=8 &
build c8l2a -DGC_STEP2_L2_AHEAD=1 -DGC_STEP2_MIN_CTAS_ALL=8 &
wait
