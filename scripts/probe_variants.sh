#!/bin/bash
# A/B of step-kernel variants on the GPU box: fixed-cost probe per variant (scripts/build_variants.sh <names> first)
cd "$(dirname "$0")/.."
run() { echo "=== $1"; shift; env "$@" python scripts/fixed_cost_probe.py 2>&1 | tail -9; }
run default
for v in "$@"; do
  echo "--- parity of $v"; GC_LIBGYMCOOK=$PWD/build/variants/libgymcook_$v.so python -m pytest tests/test_env_gpu.py -m gpu -x -q -k "plain or plans" 2>&1 | tail -2
  run $v GC_LIBGYMCOOK=$PWD/build/variants/libgymcook_$v.so
done
