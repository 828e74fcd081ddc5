"""cfg-2 step timing only (scratch helper)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from quick_time import time_step
for _ in range(5):
    time_step("partial-divider_tl", 2, 1 << 20)
time_step("partial-divider_tl", 2, 1 << 24, ring=2, iters=60)
