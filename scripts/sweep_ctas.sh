#!/bin/bash
# sweep the persistent-grid depth of the step kernel (scratch helper)
for c in 2 3 4 5 6 8; do echo "== GC_STEP_CTAS_PER_SM=$c"; GC_STEP_CTAS_PER_SM=$c python scripts/quick_time.py 2>&1 | head -2; done
