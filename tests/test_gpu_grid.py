"""k_grid (overview compositor) through the C ABI: bit-exact against overview images composed by the live
reference and against the oracle on random panel assignments."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def test_golden_overviews(built):
    import torch
    from reasoning_image_generation_b200.compose import GridCompositor, GridLayout
    g = np.load(os.path.join(GOLDEN, 'grid_panels.npz'))
    pix = np.load(os.path.join(GOLDEN, 'seq_panels.npz'))['pixels']
    dev = torch.from_numpy(np.ascontiguousarray(pix)).cuda()
    for case, n_states in ((0, 3), (1, 5), (2, 3)):
        L = GridLayout(512, 512, n_states=n_states, num_options=4)
        c = GridCompositor(L, max_samples=4)
        got = c.compose(dev.data_ptr(), g[f'index_{case}'][None, :])
        assert (got[0] == g[f'grid_{case}']).all(), case
        c.close()


def test_random_assignments_blank_cells_and_device_output(built):
    import torch
    from oracle import grid_oracle
    from reasoning_image_generation_b200.compose import GridCompositor, GridLayout
    rng = np.random.RandomState(4)
    panels = np.full((40, 512, 512, 3), 255, np.uint8)
    for p in panels:
        for _ in range(12):
            x, y = rng.randint(0, 480), rng.randint(0, 480)
            p[y:y + rng.randint(2, 120), x:x + rng.randint(2, 120)] = rng.randint(0, 256, 3)
    panels[7] = rng.randint(0, 256, (512, 512, 3))          # noise: every rounding case
    dev = torch.from_numpy(panels).cuda()
    L = GridLayout(512, 512, n_states=3, num_options=4)
    c = GridCompositor(L, max_samples=64)
    idx = rng.randint(0, 40, (50, 7)).astype(np.int32)
    idx[3, 2] = -1
    idx[10, :] = 7
    out = torch.zeros((50, L.grid_h, L.grid_w, 3), dtype=torch.uint8, device='cuda')
    c.set_profiling(True)
    c.compose_to_device(dev.data_ptr(), idx, out.data_ptr())
    c.sync()
    got = out.cpu().numpy()
    assert (got == c.compose(dev.data_ptr(), idx)).all()
    white = np.full((512, 512, 3), 255, np.uint8)
    for s in (0, 3, 10, 49):
        imgs = [panels[i] if i >= 0 else white for i in idx[s]]
        ref = grid_oracle.compose(512, 512, imgs[:3], imgs[3:])
        assert (got[s] == ref).all(), s
    st = c.stats()
    assert st['timed_samples'] == 50 and st['compose_ms'] > 0
    c.close()


def test_rgb_panels_and_errors(built):
    import torch
    from reasoning_image_generation_b200.compose import GridCompositor, GridError, GridLayout
    L = GridLayout(512, 512, n_states=3, num_options=4, bg_color=(250, 240, 230))
    rng = np.random.RandomState(2)
    panels = rng.randint(0, 256, (7, 512, 512, 3)).astype(np.uint8)
    bgr = torch.from_numpy(panels).cuda()
    rgb = torch.from_numpy(np.ascontiguousarray(panels[..., ::-1])).cuda()
    a = GridCompositor(L, max_samples=2, channel_order='bgr')
    b = GridCompositor(L, max_samples=2, channel_order='rgb')
    idx = np.arange(7, dtype=np.int32)[None]
    assert (a.compose(bgr.data_ptr(), idx)[..., ::-1] == b.compose(rgb.data_ptr(), idx)).all()
    with pytest.raises(GridError):
        a.compose(bgr.data_ptr(), np.zeros((3, 7), np.int32))
    a.close()
    b.close()
