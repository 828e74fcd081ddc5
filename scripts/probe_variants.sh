#!/bin/bash
# A/B of step-kernel variants on the GPU box: fixed-cost probe per variant (scripts/build_variants.sh first)
cd "$(dirname "$0")/.."
run() { echo "=== $1"; shift; env "$@" python scripts/fixed_cost_probe.py 2>&1 | tail -9; }
run default
run one_table_copy GC_STEP_TABLE_COPIES=1
run no_pdl GC_STEP_NO_PDL=1
for v in c7 c8 t128 l2a c8l2a; do
  run $v GC_LIBGYMCOOK=$PWD/build/variants/libgymcook_$v.so
done
