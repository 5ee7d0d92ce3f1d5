"""GPU hash + all-pairs Hamming kernels against the numpy oracle (oracle/dedup_oracle.py).  Parity with the
reference is unpinned for this stage (the reference has no such filter); these tests pin CUDA == oracle."""
import numpy as np
import pytest

from reasoning_image_generation_b200 import Renderer
from reasoning_image_generation_b200.dedup import DedupFilter, row_blocks
from reasoning_image_generation_b200.synth import synthetic_scene
from oracle import dedup_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ctx(built):
    import torch
    r = Renderer(512, 512, device=0, max_panels=4096, max_elements=16384)
    yield r, DedupFilter(r), torch
    r.close()


def test_hash_kernel_matches_oracle_on_rendered_panels(ctx):
    r, f, torch = ctx
    sc = synthetic_scene(300, seed=21)
    for layout in ('bgr', 'rgb'):
        imgs = torch.empty((300, 512, 512, 3), dtype=torch.uint8, device='cuda')
        r.render_to_device(sc, imgs.data_ptr(), layout); r.sync()
        h = f.hash_images(imgs, layout); torch.cuda.synchronize(); r.sync()
        got = h.cpu().numpy().view(np.uint64)
        host = imgs.cpu().numpy()
        want = np.stack([dedup_oracle.phash(host[i], layout) for i in range(300)])
        assert (got == want).all()


def test_hash_kernel_on_noise(ctx):
    r, f, torch = ctx
    g = torch.Generator(device='cuda'); g.manual_seed(5)
    imgs = torch.randint(0, 256, (64, 512, 512, 3), dtype=torch.uint8, device='cuda', generator=g)
    h = f.hash_images(imgs, 'rgb'); torch.cuda.synchronize(); r.sync()
    host = imgs.cpu().numpy()
    want = np.stack([dedup_oracle.phash(host[i], 'rgb') for i in range(64)])
    assert (h.cpu().numpy().view(np.uint64) == want).all()


@pytest.mark.parametrize('n,thr', [(1, 10), (257, 10), (5000, 10), (5000, 0), (3001, 40), (4096, 3), (4097, 3), (20011, 10)])
def test_hamming_scan_matches_oracle(ctx, n, thr):
    r, f, torch = ctx
    rng = np.random.default_rng(n + thr)
    h = rng.integers(-2 ** 63, 2 ** 63 - 1, (n, 4), dtype=np.int64)
    # plant near duplicates at assorted distances, including exact copies and chains
    for k in range(n // 50):
        i, j = sorted(rng.integers(0, n, 2))
        if i == j:
            continue
        flips = int(rng.integers(0, thr + 3))
        v = h[i].copy().view(np.uint64)
        for b in rng.choice(256, flips, replace=False):
            v[b // 64] ^= np.uint64(1) << np.uint64(b % 64)
        h[j] = v.view(np.int64)
    # pairs that pass the 64-bit pre-filter of k_hamming (equal first words) but differ by thr + 1 bits further up, and
    # pairs that differ only in the first words
    for k in range(n // 200):
        i, j = sorted(rng.integers(0, n, 2))
        if i == j:
            continue
        v = h[i].copy().view(np.uint64)
        for b in rng.choice(np.arange(64, 256), thr + 1 - (k & 1), replace=False):
            v[b // 64] ^= np.uint64(1) << np.uint64(b % 64)
        h[j] = v.view(np.int64)
    hu = h.view(np.uint64)
    want = dedup_oracle.drop_flags(hu, thr)
    hd = torch.from_numpy(h).cuda()
    drop, hits = f.scan(hd, thr); torch.cuda.synchronize(); r.sync()
    assert (drop.cpu().numpy() == want).all()
    assert int(hits.item()) == dedup_oracle.count_pairs(hu, thr)
    # row blocks as the ranks would scan them
    u = torch.zeros(n, dtype=torch.uint8, device='cuda')
    for b, e in row_blocks(n, 3):
        f.scan(hd, thr, b, e, drop=u)
    torch.cuda.synchronize(); r.sync()
    assert (u.cpu().numpy() == want).all()


def test_render_hash_dedup_pipeline_drops_planted_duplicates(ctx):
    r, f, torch = ctx
    sc = synthetic_scene(400, seed=33)
    # panels 400..419 repeat panels 0..19 exactly: they must be dropped, the originals kept
    from reasoning_image_generation_b200.scene import SceneBuilder
    from reasoning_image_generation_b200.synth import scene_to_elements
    sb = SceneBuilder(512, 512)
    for p in list(range(400)) + list(range(20)):
        els, ug = scene_to_elements(sc, p)
        sb.add_panel(els, ug)
    big = sb.finish()
    imgs = torch.empty((420, 512, 512, 3), dtype=torch.uint8, device='cuda')
    r.render_to_device(big, imgs.data_ptr(), 'bgr'); r.sync()
    h = f.hash_images(imgs, 'bgr')
    drop, hits = f.scan(h, 0); torch.cuda.synchronize(); r.sync()
    d = drop.cpu().numpy()
    assert d[400:].all()
    want = dedup_oracle.drop_flags(h.cpu().numpy().view(np.uint64), 0)
    assert (d == want).all()


def test_fused_hash_equals_standalone_hash_and_oracle(ctx):
    """rig_render_resident_hashed: block sums accumulated inside the rasteriser == hashing the written pixels."""
    r, f, torch = ctx
    from reasoning_image_generation_b200.scene import SceneBuilder
    sc = synthetic_scene(700, seed=44)
    for layout in ('rgb', 'bgr'):
        imgs = torch.empty((700, 512, 512, 3), dtype=torch.uint8, device='cuda')
        fused = torch.empty((700, 4), dtype=torch.int64, device='cuda')
        r.upload(sc, layout)
        r.render_resident_hashed(imgs.data_ptr(), fused.data_ptr()); r.sync()
        plain = torch.empty_like(imgs)
        r.render_resident(plain.data_ptr()); r.sync()
        assert torch.equal(imgs, plain)                      # the hashed variant writes the same pixels
        sep = f.hash_images(imgs, layout); torch.cuda.synchronize(); r.sync()
        assert torch.equal(fused, sep)
        host = imgs[:40].cpu().numpy()
        want = np.stack([dedup_oracle.phash(host[i], layout) for i in range(40)])
        assert (fused[:40].cpu().numpy().view(np.uint64) == want).all()
    # a non-white background and an all-empty batch
    sb = SceneBuilder(512, 512, bg_color=(12, 200, 77))
    for p in range(6):
        sb.add_panel([], p % 2 == 0)
    sb.add_panel([dict(kind='heart', size=150, fill=True, stroke_width=3, center=(256, 256), angle=30, color=(250, 250, 250))], True)
    s2 = sb.finish()
    imgs = torch.empty((7, 512, 512, 3), dtype=torch.uint8, device='cuda')
    fused = torch.empty((7, 4), dtype=torch.int64, device='cuda')
    r.upload(s2, 'rgb'); r.render_resident_hashed(imgs.data_ptr(), fused.data_ptr()); r.sync()
    host = imgs.cpu().numpy()
    want = np.stack([dedup_oracle.phash(host[i], 'rgb') for i in range(7)])
    assert (fused.cpu().numpy().view(np.uint64) == want).all()
