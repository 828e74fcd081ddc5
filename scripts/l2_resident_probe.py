"""Is the step kernel's steady state bound by DRAM?  The same launches over ONE batch (16 MB of state + 2 MB of
actions + 1 MB of results: resident in the 126 MB L2) against the ring of 16 batches that bench.py uses."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fixed_cost_probe as P
for k in (4, 8):
    n = P.TILE * k
    print("tiles/CTA %d: ring of 16 batches %.2f us per launch, one L2-resident batch %.2f us" % (
        k, P.per_launch_us(n, ring=16), P.per_launch_us(n, ring=1)))
