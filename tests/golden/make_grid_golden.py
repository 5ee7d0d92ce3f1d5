"""Freezes overview images composed by the LIVE reference (imports /root/reference/src/layout.py; run in the build
container only):  python tests/golden/make_grid_golden.py  ->  tests/golden/grid_panels.npz"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, '/root/reference/src')
import layout as ref_layout  # noqa: E402

pix = np.load(os.path.join(HERE, 'seq_panels.npz'))['pixels']      # panels rendered by the live reference (BGR)
out = {}
with tempfile.TemporaryDirectory() as d:
    k = 0
    for case, (n_states, n_opt) in enumerate(((3, 4), (5, 4), (3, 4))):
        states = [{'state_img': pix[(k + i) % len(pix)], 'state_path': None, 'proto_path': None} for i in range(n_states)]
        k += n_states
        cands = [{'img': pix[(k + i) % len(pix)], 'path': None, 'is_correct': i == 1} for i in range(n_opt)]
        k += n_opt
        grid, cells_meta, *_rest, grid_h, cell = ref_layout.compose_grid(512, 512, states, cands, d, n_opt)
        out[f'grid_{case}'] = grid
        out[f'index_{case}'] = np.array([(k - n_states - n_opt + i) % len(pix) for i in range(n_states + n_opt)], dtype=np.int32)
        out[f'bbox_{case}'] = np.array([c['bbox'] for c in cells_meta], dtype=np.int32)
np.savez_compressed(os.path.join(HERE, 'grid_panels.npz'), **out)
print({k: v.shape for k, v in out.items()})
