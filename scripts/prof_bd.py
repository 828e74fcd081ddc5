"""One shape of gc_bd_posterior for ncu (scratch helper)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from time_planners import posterior
posterior(1 << 20, 8, 8, 5, 2)
