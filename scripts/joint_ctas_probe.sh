#!/bin/bash
# joint solver: resident searches per SM (arena count) vs throughput on cfg-3 states
cd "$(dirname "$0")/.."
for cfg in "8 128" "12 64" "16 64"; do set -- $cfg; echo "== GC_JOINT_CTAS_PER_SM=$1 GC_JOINT_THREADS=$2"; GC_JOINT_CTAS_PER_SM=$1 GC_JOINT_THREADS=$2 python scripts/cfg3_probe.py 2>&1 | grep -E "n=2\^(18)|rror"; done
