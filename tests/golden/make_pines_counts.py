"""Derive the 40x40 pines bin-count fixture from the reference's data file.

Run in the build container only (reads /root/reference/finpines.csv, which does not exist on
the GPU box).  Binning rule follows cox_process_utils.py:28-55 (floor(x*n), clamp upper edge).
Writes mfm_b200/data/pines_counts_40x40.txt (1600 small integers, row-major [row, col]).
"""
import sys
import numpy as np

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/finpines.csv"
pts = np.genfromtxt(src, delimiter=",")
assert pts.ndim == 2 and pts.shape[1] == 2, pts.shape
n = 40
counts = np.zeros((n, n), dtype=np.int64)
for px, py in pts * n:
    r, c = int(np.floor(px)), int(np.floor(py))
    r -= r == n
    c -= c == n
    counts[r, c] += 1
out = "mfm_b200/data/pines_counts_40x40.txt"
np.savetxt(out, counts, fmt="%d")
print(out, "sum", counts.sum(), "max", counts.max(), "nonzero", (counts > 0).sum())
