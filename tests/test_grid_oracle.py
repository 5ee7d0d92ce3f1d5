"""Overview compositor, CPU side: the oracle against overview images composed by the live reference (tests/golden/
grid_panels.npz, made by tests/golden/make_grid_golden.py), the numpy restatement of cv2's INTER_AREA against the installed
cv2, and the host layout logic (overlay + cell map) of the product."""
import os

import cv2
import numpy as np
import pytest

from conftest import GOLDEN
from oracle import grid_oracle
from reasoning_image_generation_b200.compose import GridLayout


@pytest.fixture(scope='module')
def gold():
    g = np.load(os.path.join(GOLDEN, 'grid_panels.npz'))
    pix = np.load(os.path.join(GOLDEN, 'seq_panels.npz'))['pixels']
    return g, pix


def cases(g):
    for case, n_states in ((0, 3), (1, 5), (2, 3)):
        yield case, n_states, g[f'grid_{case}'], g[f'index_{case}'], g[f'bbox_{case}']


def test_oracle_equals_the_live_reference(gold):
    g, pix = gold
    for case, n_states, grid, index, bbox in cases(g):
        got = grid_oracle.compose(512, 512, [pix[i] for i in index[:n_states]], [pix[i] for i in index[n_states:]])
        assert got.shape == grid.shape and (got == grid).all(), case


def test_numpy_area_resize_equals_cv2():
    rng = np.random.RandomState(1)
    for s, d in ((512, 118), (512, 78), (200, 67)):
        src = rng.randint(0, 256, (s, s, 3)).astype(np.uint8)
        src[s // 5:s // 2, s // 10:s - 20] = 255
        assert (grid_oracle.resize_area(src, d, d) == cv2.resize(src, (d, d), interpolation=cv2.INTER_AREA)).all(), (s, d)


def test_layout_overlay_and_cell_map_reproduce_the_reference(gold):
    """Host logic of the product: constant overlay + per-pixel cell map; filling the mapped pixels with cv2's own resize
    must give the reference's image, and the cell boxes must be the reference's cells_meta boxes."""
    g, pix = gold
    for case, n_states, grid, index, bbox in cases(g):
        L = GridLayout(512, 512, n_states=n_states, num_options=4)
        assert (L.grid_h, L.grid_w) == grid.shape[:2]
        assert [m['bbox'] for m in L.cells_meta] == bbox.tolist()
        out = L.overlay.copy()
        for k, (cx, cy) in enumerate(L.cells):
            patch = cv2.resize(pix[index[k]], (L.cell_size, L.cell_size), interpolation=cv2.INTER_AREA)
            sel = L.cell_of_pixel[cy:cy + L.cell_size, cx:cx + L.cell_size] == k
            out[cy:cy + L.cell_size, cx:cx + L.cell_size][sel] = patch[sel]
        assert (out == grid).all(), case
        # borders belong to the overlay: the outer ring of every cell is not a panel pixel
        cx, cy = L.cells[0]
        assert (L.cell_of_pixel[cy, cx:cx + L.cell_size] == -1).all() and (L.cell_of_pixel[cy + 1, cx + 1:cx + L.cell_size - 1] == 0).all()


def test_layout_without_borders_or_labels():
    L = GridLayout(512, 512, n_states=3, num_options=4, show_labels=False, show_border=False)
    cx, cy = L.cells[0]
    assert (L.cell_of_pixel[cy:cy + L.cell_size, cx:cx + L.cell_size] == 0).all()
    assert (L.overlay[L.cell_of_pixel < 0][:, 0] <= 255).all()


def test_grid_symbols_exported(built):
    import re
    from conftest import ROOT
    from reasoning_image_generation_b200 import _abi, compose
    lib = _abi.load()
    txt = open(os.path.join(ROOT, 'include', 'rig_grid.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    names = sorted(set(re.findall(r'\b(rig_grid_[a-z_]+)\s*\(', txt)))
    assert len(names) == 7 and sorted(compose.GRID_EXPORTS) == names
    for n in names:
        assert hasattr(lib, n), n
