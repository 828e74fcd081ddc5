#!/bin/bash
# joint solver: per-search state budget (GC_JOINT_BUDGET) vs failures and time
cd "$(dirname "$0")/.."
for b in 98304 49152 24576; do
  echo "== GC_JOINT_BUDGET=$b"
  GC_JOINT_BUDGET=$b python scripts/time_joint.py
  GC_JOINT_BUDGET=$b GC_STATUS=1 timeout 300 python scripts/time_cfg5.py 1024 60 partial-divider_tl,partial-divider_tl 2>&1 | grep -v single
done
