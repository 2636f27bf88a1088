#!/usr/bin/env python
"""TEST INFRASTRUCTURE — extracts 16 of the 248 Dragon Lake Parking cases of the reference's data/dlp.data
(first 8 start candidates each) into tests/golden/dlp_cases.npz as plain arrays, using the product's own
reader (hope_b200/dlp.py).  Runs in the build container only; the GPU box has no /root/reference."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hope_b200 import dlp  # noqa: E402

SEL = [0, 7, 31, 50, 77, 100, 123, 150, 171, 190, 205, 220, 233, 240, 245, 247]


def main(ref="/root/reference"):
    cases = dlp.read_dlp(os.path.join(ref, "data", "dlp.data"))
    fx = {"case_ids": np.array(SEL)}
    for j, c in enumerate(SEL):
        rings = cases[c]["rings"]
        pad = np.zeros((len(rings), 4, 2))
        for k, x in enumerate(rings):
            pad[k, :len(x)] = x
        fx[f"starts_{j}"] = cases[c]["starts"][:8]
        fx[f"dest_{j}"] = cases[c]["dest"]
        fx[f"ring_nv_{j}"] = np.array([len(x) for x in rings], dtype=np.int32)
        fx[f"rings_{j}"] = pad
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "dlp_cases.npz"), **fx)


if __name__ == "__main__":
    main()
