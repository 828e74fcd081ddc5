#!/bin/bash
# e2e A/B runs of the host-driven step (scripts/e2e_breakdown.py, first four lines = the public calls)
cd "$(dirname "$0")/.."
for c in 1 2 3 4; do echo "== GC_E2E_CHUNKS=$c"; GC_E2E_CHUNKS=$c python scripts/e2e_breakdown.py 2>&1 | head -4; done
echo "== zero-copy results (GC_E2E_ZEROCOPY=1)"; GC_E2E_ZEROCOPY=1 python scripts/e2e_breakdown.py 2>&1 | head -4
