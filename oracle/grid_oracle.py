"""oracle/grid_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's overview compositor:
    compose_grid / fit_into_cell / make_query_image        /root/reference/src/layout.py:17-192
`compose` repeats compose_grid's cv2 calls in its order (the installed cv2 wheel is the third-party arithmetic, as for the
panel renderer); `resize_area` restates cv2.resize(INTER_AREA) for non-integer scales (OpenCV imgproc/src/resize.cpp:
computeResizeAreaTab + ResizeArea_Invoker, float32 work type) in numpy.
Parity status: PINNED -- tests/golden/grid_panels.npz holds overview images composed by the live reference
(tests/golden/make_grid_golden.py imports /root/reference/src/layout.py in the build container), and `resize_area` is
checked against the installed cv2 on random images.  Only tests/ and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import math

import numpy as np


def area_tab(ssize: int, dsize: int):
    """computeResizeAreaTab: list of (dst index, src index, float32 weight) in OpenCV's order."""
    scale = ssize / dsize
    tab = []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1, sx2 = math.ceil(fsx1), math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((dx, sx1 - 1, np.float32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab.append((dx, sx, np.float32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((dx, sx2, np.float32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return tab


def resize_area(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(src, (dw, dh), interpolation=INTER_AREA) for scales that are not whole numbers (uint8, 3 channels)."""
    sh, sw = src.shape[:2]
    xt, yt = area_tab(sw, dw), area_tab(sh, dh)
    dst = np.zeros((dh, dw, 3), np.uint8)
    acc = np.zeros((dw, 3), np.float32)
    prev = yt[0][0]

    def store(row, v):
        dst[row] = np.clip(np.rint(v), 0, 255).astype(np.uint8)
    for dy, sy, beta in yt:
        buf = np.zeros((dw, 3), np.float32)
        S = src[sy].astype(np.float32)
        for dx, sx, alpha in xt:
            buf[dx] = buf[dx] + S[sx] * alpha
        if dy != prev:
            store(prev, acc)
            acc = beta * buf
            prev = dy
        else:
            acc = acc + beta * buf
    store(prev, acc)
    return dst


def compose(W, H, state_imgs, option_imgs, margin=20, padding_v=20, show_labels=True, show_border=True, bg_color=(255, 255, 255)):
    """compose_grid's image (BGR uint8[grid_h, W, 3]) from the state panels and the option panels."""
    import cv2
    num_options = len(option_imgs)
    cols_seq, cols_opt = len(state_imgs) + 1, num_options
    max_cell_w = (W - 2 * margin) // max(1, max(cols_seq, cols_opt))
    max_cell_h = (H - 2 * margin - padding_v) // 2
    cell = max(1, min(max_cell_w, max_cell_h))
    grid_h = 2 * cell + padding_v + 2 * margin
    grid = np.full((grid_h, W, 3), (bg_color[2], bg_color[1], bg_color[0]), dtype=np.uint8)
    font = cv2.FONT_HERSHEY_SIMPLEX
    font_scale = max(0.25, min(0.8, cell / 240.0))

    def fit(src):
        Hs, Ws = src.shape[:2]
        scale = min(cell / Ws, cell / Hs)
        nw, nh = max(1, int(round(Ws * scale))), max(1, int(round(Hs * scale)))
        r = cv2.resize(src, (nw, nh), interpolation=cv2.INTER_AREA if scale < 1.0 else cv2.INTER_CUBIC)
        patch = np.full((cell, cell, 3), 255, dtype=np.uint8)
        ox, oy = (cell - nw) // 2, (cell - nh) // 2
        patch[oy:oy + nh, ox:ox + nw] = r
        return patch

    def query():
        im = np.full((cell, cell, 3), 255, dtype=np.uint8)
        thickness = max(1, cell // 80)
        fs = cell / 100.0
        (tw, th), _ = cv2.getTextSize('?', font, fs, thickness)
        if tw > cell * 0.8:
            fs = fs * (cell * 0.8 / tw)
            (tw, th), _ = cv2.getTextSize('?', font, fs, thickness)
        cv2.putText(im, '?', ((cell - tw) // 2, (cell + th) // 2), font, fs, (0, 0, 0), thickness, lineType=cv2.LINE_AA)
        return im

    def place(patch, x, y, label):
        grid[y:y + cell, x:x + cell] = patch
        if show_border:
            cv2.rectangle(grid, (x, y), (x + cell - 1, y + cell - 1), (0, 0, 0), 1)
        if show_labels and label:
            cv2.putText(grid, label, (x + 3, y + cell + int(12 * font_scale) + 6), font, font_scale, (0, 0, 0), 1, cv2.LINE_AA)
    sx, top = (W - cols_seq * cell) // 2, margin
    for i in range(cols_seq):
        place(fit(state_imgs[i]) if i < len(state_imgs) else query(), sx + i * cell, top, f'S{i}')
    ox, bottom = (W - cols_opt * cell) // 2, margin + cell + padding_v
    for i in range(cols_opt):
        place(fit(option_imgs[i]), ox + i * cell, bottom, chr(65 + i))
    return grid
