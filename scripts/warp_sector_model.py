#!/usr/bin/env python
"""Sectors a warp-wide 32-bit tap load of the crop-warp kernel touches, for two lane mappings,
on the bench.py box distribution (480x640x3 sources -> 256x192 crops, no rotation).

  quad   : a lane owns 4 consecutive output pixels (the shipped kernel): lane L of a warp reads
           the taps of output column 4 L + j in load j
  column : a lane owns 1 output column (the round-2 lead in DESIGN.md section 9)

The model counts distinct 32-byte sectors over the 32 lanes' 4-byte words; it is checked against
the measured l1tex__t_sectors / l1tex__t_requests of the shipped kernel (17.1, profiles/
r01n_warp_ncu_busiest.txt and DESIGN.md) before it is used to predict the other mapping.

    python scripts/warp_sector_model.py [crops]
"""
import sys

import numpy as np


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    ws, hs, dw, dh = 640, 480, 192, 256
    rng = np.random.RandomState(0)
    bw, bh = rng.uniform(40, 400, n), rng.uniform(60, 440, n)
    bx, by = rng.uniform(0, 1, n) * (ws - bw), rng.uniform(0, 1, n) * (hs - bh)
    asp = dw / dh
    sw = np.where(bw > asp * bh, bw, bh * asp) * 1.25          # source pixels across the crop
    scale = sw / dw
    x_src0 = bx + bw / 2 - sw / 2
    tot = {"quad": [0, 0], "column": [0, 0]}
    for i in range(n):
        sx = np.floor(x_src0[i] + scale[i] * np.arange(dw)).astype(np.int64)   # left tap column
        inside = (sx >= 1) & (sx < ws - 4)
        byte = sx * 3
        word = byte // 4                                        # first of the 3 aligned words
        for name in tot:
            if name == "quad":      # warp = 32 quads = 128 columns; load j: columns 4 L + j
                groups = [np.arange(g, min(g + 128, dw), 4) + j
                          for g in range(0, dw, 128) for j in range(4)]
            else:                   # warp = 32 consecutive columns
                groups = [np.arange(g, g + 32) for g in range(0, dw, 32)]
            for cols in groups:
                cols = cols[cols < dw]
                cols = cols[inside[cols]]
                if len(cols) == 0:
                    continue
                for k in range(3):  # the three aligned words of a run are three loads
                    sectors = np.unique((word[cols] + k) * 4 // 32)
                    tot[name][0] += len(sectors)
                    tot[name][1] += 1
    for name, (s, r) in tot.items():
        print(f"{name:7s}: {s / r:5.2f} sectors per warp-wide tap load ({r} loads modelled)")


if __name__ == "__main__":
    main()
