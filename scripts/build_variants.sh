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
for v in "$@"; do
  case $v in
    c7) build c7 -DGC_STEP2_MIN_CTAS_ALL=7 & ;;
    c8) build c8 -DGC_STEP2_MIN_CTAS_ALL=8 & ;;
    t128) build t128 -DGC_STEP2_THREADS=128 -DGC_STEP2_MIN_CTAS_ALL=12 & ;;
    nol2a) build nol2a -DGC_STEP2_L2_AHEAD=0 & ;;
    ilp2c4) build ilp2c4 -DGC_STEP2_ILP=2 -DGC_STEP2_MIN_CTAS_ALL=4 & ;;
    ilp2c5) build ilp2c5 -DGC_STEP2_ILP=2 -DGC_STEP2_MIN_CTAS_ALL=5 & ;;
    ilp2t128) build ilp2t128 -DGC_STEP2_ILP=2 -DGC_STEP2_THREADS=128 -DGC_STEP2_MIN_CTAS_ALL=8 & ;;
    jpe) build jpe -DGC_JOINT_PER_ENTRY & ;;
    oldpol) build oldpol -DGC_JOINT_TREE_STATES=98304 -DGC_JOINT_WIDEN_STATES=98304 & ;;  # + GC_JOINT_UCS_FALLBACK=1 at run time
    t32w32) build t32w32 -DGC_JOINT_TREE_STATES=32768 -DGC_JOINT_WIDEN_STATES=32768 & ;;
    t64w16) build t64w16 -DGC_JOINT_TREE_STATES=65536 -DGC_JOINT_WIDEN_STATES=16384 & ;;
    t16w8) build t16w8 -DGC_JOINT_TREE_STATES=16384 -DGC_JOINT_WIDEN_STATES=8192 & ;;
    nb100) build nb100 -DGC_SEARCH_NODE_BUDGET=100000 & ;;
    nb25) build nb25 -DGC_SEARCH_NODE_BUDGET=25000 & ;;
    seeded) build seeded -DGC_JOINT_SEEDED & ;;
    t16w2) build t16w2 -DGC_JOINT_TREE_STATES=16384 -DGC_JOINT_WIDEN_STATES=2048 & ;;
  esac
done
wait
