"""Device timing of gc_bd_posterior only (scratch helper)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from time_planners import posterior

if __name__ == "__main__":
    posterior(1 << 18, 8, 8, 5, 2)
    posterior(1 << 20, 8, 8, 5, 2)
    posterior(1 << 22, 8, 8, 5, 2)
    posterior(1 << 18, 36, 15, 25, 3)
    posterior(1 << 18, 84, 24, 25, 4)
    posterior(1 << 20, 8, 8, 5, 2, torch.float64)
