"""Synthetic scene descriptions of the sequence-panel shape, vectorised in numpy.

Distributions follow what the reference's sampler and rules emit (/root/reference/src/sample.py:46-222,
/root/reference/src/rules.py:22-334): elements per panel from the measured histogram, kind uniform over the
11 kinds, fill P=2/3, stroke U{1,2,3}, snap angles {0,45,90,135,180} plus rule deltas, grid panels (size 102,
cell-centred) half of the time, otherwise size 128+U{-42..42} at a uniform in-canvas centre, and ~4 % of
elements pushed across an edge (the translate rule does not clamp) so wrap-around copies are exercised.
Seeded: the same (n_panels, seed) always gives the same scene.
"""
from __future__ import annotations

import numpy as np

from .scene import Scene

_NEL_P = np.array([0.02, 0.41, 0.38, 0.13, 0.034, 0.013, 0.010, 0.002, 0.001])


def synthetic_scene(n_panels: int, seed: int = 0, W: int = 512, H: int = 512) -> Scene:
    rng = np.random.default_rng(seed)
    nel = rng.choice(9, size=n_panels, p=_NEL_P / _NEL_P.sum()).astype(np.int64)
    begin = np.zeros(n_panels + 1, dtype=np.uint32)
    begin[1:] = np.cumsum(nel)
    ne = int(begin[-1])
    use_grid = rng.random(n_panels) < 0.5
    panel_of = np.repeat(np.arange(n_panels), nel)
    g = use_grid[panel_of]
    kind = rng.integers(0, 11, ne).astype(np.uint8)
    fill = (rng.random(ne) < 2.0 / 3.0).astype(np.uint8)
    stroke = rng.integers(1, 4, ne).astype(np.int32)
    angle = rng.choice(np.array([0, 45, 90, 135, 180]), ne)
    delta = rng.choice(np.array([0, 0, 0, 30, 45, 60, 90]), ne) * rng.integers(0, 4, ne)
    angle = ((angle + delta) % 360).astype(np.int32)
    angle[kind == 1] = 0
    size = np.where(g, int(W * 0.2), 128 + rng.integers(-42, 43, ne)).astype(np.int32)
    half = size // 2 + 5
    cx = half + rng.random(ne) * np.maximum(W - 2 * half, 1)
    cy = half + rng.random(ne) * np.maximum(H - 2 * half, 1)
    cx = np.floor(cx)
    cy = np.floor(cy)
    # grid panels: any point of a cell snaps to the cell centre inside the library
    cell = rng.integers(0, 9, ne)
    cx = np.where(g, (cell % 3 + 0.5) * W / 3.0 + rng.integers(-20, 21, ne), cx)
    cy = np.where(g, (cell // 3 + 0.5) * H / 3.0 + rng.integers(-20, 21, ne), cy)
    # un-clamped translations: 4 % of the free elements leave the canvas partly or completely
    moved = (~g) & (rng.random(ne) < 0.04)
    step = rng.choice(np.array([-2, -1, 1, 2]), ne) * 170
    axis = rng.random(ne) < 0.5
    cx = np.where(moved & axis, cx + step, cx)
    cy = np.where(moved & ~axis, cy + step, cy)
    rgb = np.floor(30 + rng.random((ne, 3)) * 190).astype(np.uint8)
    return Scene(W, H, (255, 255, 255), 3, begin, use_grid.astype(np.uint8), kind, fill, np.zeros(ne, np.uint8), stroke,
                 size, angle, cx.astype(np.float64), cy.astype(np.float64), rgb.reshape(-1))


def scene_to_elements(scene: Scene, panel: int):
    """Panel `panel` of a Scene as reference-style element dicts (for the oracle / CPU baseline)."""
    from .scene import KINDS
    e0, e1 = int(scene.begin[panel]), int(scene.begin[panel + 1])
    flip_names = {0: None, 1: 'horizontal', 2: 'vertical', 3: 'both'}
    els = []
    for i in range(e0, e1):
        els.append(dict(kind=KINDS[int(scene.kind[i])], size=int(scene.size[i]), fill=bool(scene.fill[i]),
                        stroke_width=int(scene.stroke[i]), center=(float(scene.cx[i]), float(scene.cy[i])),
                        angle=int(scene.angle[i]), flip_mode=flip_names[int(scene.flip[i])],
                        color=(int(scene.rgb[3 * i]), int(scene.rgb[3 * i + 1]), int(scene.rgb[3 * i + 2]))))
    return els, bool(scene.flags[panel] & 1)
