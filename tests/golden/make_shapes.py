#!/usr/bin/env python
"""Regenerate tests/golden/shapes.npz — the target-shape library the assembly env trains on.

TEST/FIXTURE INFRASTRUCTURE.  The reference loads this data from `fig/results.pkl`
(assembly.py:112-119), a blob that is missing from the reference checkout
(.MISSING_LARGE_BLOBS) but is a pure function of `fig/*.png`: it is produced at import time by
marl_llm/cfg/assembly_cfg.py:32-149.  This script restates that preprocessing with cv2 + numpy
only (the reference also drags in matplotlib, whose single contribution is
`imshow(..., origin='lower').get_extent()` == (-0.5, W-0.5, -0.5, H-0.5), assembly_cfg.py:106-108).

Run in the build container (needs /root/reference/fig):
    python tests/golden/make_shapes.py
Outputs (committed): tests/golden/shapes.npz with
    l_cell[S] f64, n_g[S] i32, grid_coords[S, NG_MAX, 2] f64 (rows >= n_g are NaN),
    shape_bound_points[S, 4] f64, image_hw[S, 2] i32, names[S]
The cropped binary bitmaps (only ever used by render(), assembly.py:723) are not stored.
"""
import glob
import os
import sys

import numpy as np

CELL_PX = 36            # assembly_cfg.py:58  grid_size
TARGET_HEIGHT = 2.2     # assembly_cfg.py:95  target_hight


def shape_from_png(path):
    """One PNG -> (l_cell, grid_coords[n_g,2], bound_points[4], (H,W)).  assembly_cfg.py:44-134."""
    import cv2

    gray = cv2.imread(path, cv2.IMREAD_GRAYSCALE)
    _, binary = cv2.threshold(gray, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)   # :45-46
    ys, xs = np.nonzero(binary == 0)                                               # :49-51
    binary = binary[ys.min():ys.max() + 1, xs.min():xs.max() + 1]                  # :52
    h, w = binary.shape
    # :55 multiplies by an anti-diagonal identity == vertical flip (values stay 0/255 exactly)
    binary = binary[::-1, :].astype(np.float64)

    centres = []
    for i in range(CELL_PX, h - CELL_PX, CELL_PX):                                 # :62-79
        for j in range(CELL_PX, w - CELL_PX, CELL_PX):
            block = binary[i:i + CELL_PX, j:j + CELL_PX]
            if np.count_nonzero(block == 0) / (CELL_PX * CELL_PX) >= 1:
                centres.append([j + CELL_PX / 2, i + CELL_PX / 2])
    centres = np.array(centres, dtype=np.float64)

    xm = np.mean(centres[:, 0])                                                    # :86-89
    ym = np.mean(centres[:, 1])
    centres[:, 0] -= xm
    centres[:, 1] -= ym
    y_min, y_max = np.min(centres[:, 1]), np.max(centres[:, 1])                    # :92-95
    h_scale = TARGET_HEIGHT / (y_max - y_min)                                      # :96-98
    grid = h_scale * centres                                                       # :99

    ext = (-0.5, w - 0.5, -0.5, h - 0.5)                                           # :106-108
    new_ext = [ext[0] - xm, ext[1] - xm, ext[2] - ym, ext[3] - ym]                 # :111-112
    bound = np.array([e * h_scale for e in new_ext])                               # :127-128
    return CELL_PX * h_scale, grid, bound, (h, w)                                  # :131-134


def build(fig_dir):
    paths = sorted(glob.glob(os.path.join(fig_dir, "*.png")),
                   key=lambda p: int(os.path.basename(p).split(".")[0]))            # :139-140
    shapes = [shape_from_png(p) for p in paths]
    ng_max = max(s[1].shape[0] for s in shapes)
    coords = np.full((len(shapes), ng_max, 2), np.nan)
    for k, s in enumerate(shapes):
        coords[k, :s[1].shape[0]] = s[1]
    return dict(
        l_cell=np.array([s[0] for s in shapes]),
        n_g=np.array([s[1].shape[0] for s in shapes], dtype=np.int32),
        grid_coords=coords,
        shape_bound_points=np.stack([s[2] for s in shapes]),
        image_hw=np.array([s[3] for s in shapes], dtype=np.int32),
        names=np.array([os.path.basename(p) for p in paths]),
    )


if __name__ == "__main__":
    fig_dir = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/fig"
    data = build(fig_dir)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shapes.npz")
    np.savez_compressed(out, **data)
    print("wrote", out, "n_g =", data["n_g"].tolist(), "l_cell =", np.round(data["l_cell"], 5).tolist())
