#!/bin/bash
# time prebuilt variants of libgymcook.so (build/variants/lib_<name>.so) with scripts/quick_step.py (scratch helper)
# usage: sweep_variants.sh name[:CTAS_PER_SM] ...
cp gym-cooking_b200/libgymcook.so /tmp/lib_keep.so
for spec in "$@"; do
  v=${spec%%:*}; c=${spec#*:}; [ "$c" = "$spec" ] && c=0
  echo "== variant $v GC_LUT_CTAS_PER_SM=$c"
  cp build/variants/lib_$v.so gym-cooking_b200/libgymcook.so
  GC_LUT_CTAS_PER_SM=$c python scripts/quick_step.py 2>&1 | tail -4
done
cp /tmp/lib_keep.so gym-cooking_b200/libgymcook.so
