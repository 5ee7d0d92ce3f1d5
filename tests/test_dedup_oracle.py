"""Duplicate logic on the host: the reference's in-sample rule (pinned by recorded reference decisions) and
the definition-level checks of the perceptual-hash oracle (parity unpinned: the reference has no such filter)."""
import json
import os

import numpy as np

from conftest import GOLDEN
from oracle import dedup_oracle
from reasoning_image_generation_b200 import dedup


def test_elements_set_matches_recorded_reference_decisions():
    g = json.load(open(os.path.join(GOLDEN, 'dedup_keys.json')))
    pool = g['pool']
    for mod in (dedup_oracle, dedup):
        for pr in g['pairs']:
            assert (mod.elements_set(pool[pr['a']]) == mod.elements_set(pool[pr['b']])) == pr['equal']
        v = g['variants']
        v[4][0]['size'] = np.int64(v[4][0]['size'])
        for pr in g['variant_pairs']:
            assert (mod.elements_set(v[pr['a']]) == mod.elements_set(v[pr['b']])) == pr['equal'], pr


def test_phash_definition():
    img = np.full((512, 512, 3), 255, np.uint8)
    assert (dedup_oracle.phash(img) == 0).all()                     # flat image: no block above the mean
    img[0:32, 0:32] = 0                                             # one dark block -> every other block is above
    h = dedup_oracle.phash(img)
    assert int(np.bitwise_count(h).sum()) == 255 and not (int(h[0]) & 1)
    img2 = img.copy(); img2[100, 100] = (250, 250, 250)
    assert dedup_oracle.hamming(dedup_oracle.phash(img2), h) == 0
    rgb = np.zeros((512, 512, 3), np.uint8); rgb[:, :256, 0] = 255    # red left half (RGB layout)
    assert (dedup_oracle.phash(rgb, 'rgb') == dedup_oracle.phash(rgb[..., ::-1].copy(), 'bgr')).all()


def test_drop_rule_and_row_blocks():
    rng = np.random.default_rng(0)
    h = rng.integers(0, 2 ** 63, (700, 4), dtype=np.int64).astype(np.uint64)
    h[300] = h[10]; h[301] = h[10] ^ np.array([7, 0, 0, 0], np.uint64)      # distance 3
    h[650] = h[300]
    d = dedup_oracle.drop_flags(h, 4, block=256)
    assert set(np.nonzero(d)[0]) == {300, 301, 650}
    assert set(np.nonzero(dedup_oracle.drop_flags(h, 2, block=256))[0]) == {300, 650}
    # union over disjoint row blocks == full scan (what the ranks compute)
    for world in (2, 3, 8):
        u = np.zeros(700, np.uint8)
        blocks = dedup.row_blocks(700, world)
        assert blocks[0][0] == 0 and blocks[-1][1] == 700 and all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        for b, e in blocks:
            u |= dedup_oracle.drop_flags(h, 4, b, e, block=256)
        assert (u == d).all()
    assert dedup_oracle.count_pairs(h, 4, block=256) == 6   # (10,300) (10,301) (300,301) (10,650) (300,650) (301,650)
