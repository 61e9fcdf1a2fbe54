"""Host-only study for DESIGN.md section 7.1: how many K_B sample slots would alternative pixel tilings store at config 4?

Same rules as the planner in csrc/affinity.cu (samples ordered by column strip then raster; per tile and strip the contiguous
run of samples within reach in rows, taken when the strip is within reach in columns; runs covered by blocks of B slots
starting at multiples of 8), generalised from the current 512 x 1 raster tiles to th x tw pixel patches.  Prints, per tile
shape and block size, the best strip count and the stored slots per pixel (the current layout stores 106)."""
import math
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle_c as oc  # noqa: E402


def stored_slots(W, H, s, th, tw, B, S, R):
    col, row = (s % W).astype(np.int64), (s // W).astype(np.int64)
    strip = col * S // W
    order = np.lexsort((s, strip))
    srow, sstrip = row[order], strip[order]
    begin = np.searchsorted(sstrip, np.arange(S + 1))
    total = 0
    for ty in range(0, H, th):
        ra, rb = ty, min(H, ty + th) - 1
        lo = [begin[k] + np.searchsorted(srow[begin[k]:begin[k + 1]], ra - R, "left") for k in range(S)]
        hi = [begin[k] + np.searchsorted(srow[begin[k]:begin[k + 1]], rb + R, "right") for k in range(S)]
        for tx in range(0, W, tw):
            ca, cb = tx, min(W, tx + tw) - 1
            prev_end, cnt = 0, 0
            for k in range(S):
                if hi[k] <= lo[k]:
                    continue
                c_lo, c_hi = k * W // S, (k + 1) * W // S - 1
                if not (ca - R <= c_hi + 1 and cb + R >= c_lo - 1):
                    continue
                start = max(lo[k] & ~7, prev_end)
                while start < hi[k]:
                    cnt += 1
                    start += B
                    prev_end = start
            total += max(cnt, 1) * B * (min(H, ty + th) - ty) * (min(W, tx + tw) - tx)
    return total / (W * H)


if __name__ == "__main__":
    W, H, p, h_loc = 3840, 2160, 1000, 40.0
    s = oc.random_sampling(W, H, p, 0)
    R = int(math.floor(h_loc * math.sqrt(25 * math.log(2)))) + 1
    print(f"config 4: {W}x{H}, p={p}, reach {R} px; stored K_B slots per pixel (dense: {((p + 63) // 64) * 64})")
    for th, tw in ((1, 512), (1, 128), (4, 128), (8, 64), (16, 32), (32, 16)):
        for B in (64, 32):
            best = min(((stored_slots(W, H, s, th, tw, B, S, R), S) for S in (1, 2, 3, 4, 6, 8, 12, 16, 24)), key=lambda t: t[0])
            print(f"  tile {th:2d} rows x {tw:3d} cols, {B}-slot blocks: {best[0]:6.1f} slots/pixel with {best[1]} strips")
