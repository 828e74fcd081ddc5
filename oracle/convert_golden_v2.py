"""One-off: rewrite the `state` arrays of the planner fixtures from the ABI-1 packing (six 16-bit object
slots) to the byte-plane packing of ABI 2 (include/gymcook.h).  A lossless re-encoding of the same
reference-generated states - nothing is recomputed; gen_golden_plan.pack_env now emits the new form
directly.  Run once from the repo root: python oracle/convert_golden_v2.py"""
import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import oracle as O  # noqa: E402


def main():
    for path in sorted(glob.glob(os.path.join(HERE, "..", "tests", "golden", "*.npz"))):
        z = dict(np.load(path, allow_pickle=False))
        if "state" not in z or "state_layout" in z:
            continue
        z["state"] = O.v1_to_v2(z["state"])
        z["state_layout"] = np.array("abi2-byte-planes")
        np.savez_compressed(path, **z)
        print("converted", os.path.basename(path), z["state"].shape)


if __name__ == "__main__":
    main()
