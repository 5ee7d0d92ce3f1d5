"""Device-side PNG writer (include/rig_png.h) against independent decoders.

The parity criterion for a lossless codec: every file is a valid PNG (signature, chunk CRCs, zlib Adler-32) that decodes --
with libpng (cv2.imdecode) and with a strict pure-zlib reader -- to exactly the pixels that were rendered, which are
themselves checked against the CPU oracles elsewhere."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _encode(imgs, layout='rgb', channels=3):
    import torch
    from reasoning_image_generation_b200.png import PngEncoder, split_files
    n, H, W, _ = imgs.shape
    t = torch.from_numpy(np.ascontiguousarray(imgs)).cuda()
    enc = PngEncoder(W, H, max_images=max(n, 1), channels=channels)
    buf, off = enc.encode(t.data_ptr(), n, layout)
    files = split_files(buf, off)
    st = enc.stats()
    enc.close()
    return files, st


def _check(files, imgs_rgb):
    import cv2
    from reasoning_image_generation_b200.png import decode_png_strict
    assert len(files) == len(imgs_rgb)
    for f, ref in zip(files, imgs_rgb):
        got = decode_png_strict(f)
        cv = cv2.imdecode(np.frombuffer(f, np.uint8), cv2.IMREAD_UNCHANGED)
        assert cv is not None
        if got.shape[2] == 4:                      # RGBA files: the colour is the image, alpha is 255 everywhere
            assert (got[:, :, 3] == 255).all() and (cv[:, :, 3] == 255).all()
            got, cv = got[:, :, :3], cv[:, :, :3]
        assert got.shape == ref.shape and (got == ref).all()
        assert cv.shape == ref.shape and (cv[:, :, ::-1] == ref).all()


def test_rendered_figures_roundtrip(built):
    import torch
    from reasoning_image_generation_b200.figure import FigureRenderer
    from reasoning_image_generation_b200.synth_figure import synthetic_figure_scene
    sc = synthetic_figure_scene(48, 512, 512, seed=9, kind='single')
    r = FigureRenderer(512, 512, max_images=64)
    imgs = r.render(sc, 'rgb')
    r.close()
    files4, _ = _encode(imgs, channels=4)
    _check(files4, imgs)
    files, st = _encode(imgs)
    _check(files, imgs)
    sizes = np.array([len(f) for f in files])
    assert sizes.max() < 64 * 1024 and sizes.mean() < 24 * 1024      # vs 786,432 raw bytes
    assert st['images'] == 48 and st['png_bytes'] == sizes.sum()


def test_sequence_panels_bgr_roundtrip(built):
    """cv2.imwrite's input is BGR (src/utils.py:18-21): the encoder swaps channels so the file holds RGB."""
    from reasoning_image_generation_b200 import Renderer
    from reasoning_image_generation_b200.synth import synthetic_scene
    sc = synthetic_scene(40, seed=77)
    r = Renderer(512, 512, max_panels=64)
    bgr = r.render(sc, 'bgr')
    r.close()
    files, _ = _encode(bgr, 'bgr')
    _check(files, bgr[..., ::-1])


def test_noise_gradients_and_flat_images(built):
    """Worst and best cases: incompressible noise (every byte a literal, pieces overflow their slots), smooth gradients,
    constant images, single pixels at chunk and row boundaries."""
    rng = np.random.RandomState(5)
    W, H = 512, 64
    imgs = np.zeros((10, H, W, 3), np.uint8)
    imgs[0] = rng.randint(0, 256, (H, W, 3))
    imgs[1] = (np.arange(W)[None, :, None] + np.arange(H)[:, None, None] * 3 + np.arange(3)[None, None, :] * 85) % 256
    imgs[2] = 255
    imgs[3] = 0
    imgs[4] = 255
    imgs[4, 0, 0] = (1, 2, 3); imgs[4, 0, 15] = (4, 5, 6); imgs[4, 0, 16] = (7, 8, 9); imgs[4, 63, 511] = (200, 100, 0)
    imgs[5] = 255
    imgs[5, :, 15::16] = 0                     # last pixel of every lane chunk
    imgs[6] = 255
    imgs[6, :, 16::16] = 0                     # first pixel of every lane chunk
    imgs[7] = rng.randint(0, 2, (H, W, 1)) * 255
    imgs[8, :, :, 0] = 143; imgs[8, :, :, 1] = 144; imgs[8, :, :, 2] = 145      # 8- / 9-bit literal boundary
    imgs[9] = rng.randint(0, 256, (H, 1, 3))   # constant rows
    files, _ = _encode(imgs)
    _check(files, imgs)
    assert len(files[2]) < 2048 and len(files[0]) > H * W * 3
    files4, _ = _encode(imgs, channels=4)      # the same as RGBA files (what fig.savefig writes)
    _check(files4, imgs)
    assert files4[2][25] == 6 and files[2][25] == 2          # IHDR colour type


@pytest.mark.parametrize('size', [(256, 256), (640, 480), (1600, 1600), (16, 3)])
def test_other_sizes(built, size):
    W, H = size
    rng = np.random.RandomState(W + H)
    imgs = np.full((3, H, W, 3), 255, np.uint8)
    for k in range(3):
        for _ in range(20):
            x, y = rng.randint(0, W), rng.randint(0, H)
            imgs[k, y:y + rng.randint(1, 9), x:x + rng.randint(1, 40)] = rng.randint(0, 256, 3)
    files, _ = _encode(imgs)
    _check(files, imgs)
    files4, _ = _encode(imgs, channels=4)
    _check(files4, imgs)


def test_empty_batch_and_capacity(built):
    import torch
    from reasoning_image_generation_b200.png import PngEncoder, PngError
    enc = PngEncoder(64, 64, max_images=4, host_capacity=1 << 12)
    buf, off = enc.encode(0, 0)
    assert len(buf) == 0 and list(off) == [0]
    noise = torch.randint(0, 256, (4, 64, 64, 3), dtype=torch.uint8, device='cuda')
    buf, off = enc.encode(noise.data_ptr(), 4)          # grows the pinned destination transparently
    assert int(off[4]) > 4 * 64 * 64 * 3
    with pytest.raises(ValueError):
        enc.encode(noise.data_ptr(), 5)
    enc.close()


def test_iter_png_is_savefig_for_a_batch(built):
    """FigureRenderer.iter_png: render + encode chunk by chunk on the device; the files decode to the oracle's pixels."""
    import cv2
    from oracle.agg_oracle import AggOracle
    from reasoning_image_generation_b200.figure import FigureRenderer
    from reasoning_image_generation_b200.png import PngEncoder
    from reasoning_image_generation_b200.synth_figure import synthetic_figure_scene
    sc = synthetic_figure_scene(50, 256, 256, seed=12, kind='multi')
    ref = AggOracle().render(sc, 256, 256)
    r = FigureRenderer(256, 256, max_images=16)
    enc = PngEncoder(256, 256, max_images=16, channels=4)      # RGBA files, as fig.savefig writes them
    seen = 0
    for i0, buf, off in r.iter_png(sc, enc, chunk=16):
        for k in range(len(off) - 1):
            img = cv2.imdecode(np.asarray(buf[int(off[k]):int(off[k + 1])]), cv2.IMREAD_UNCHANGED)
            assert img.shape == (256, 256, 4) and (img[:, :, 3] == 255).all()
            assert (img[:, :, 2::-1] == ref[i0 + k]).all()
            seen += 1
    assert seen == 50
    r.close()
    enc.close()


def test_sequence_iter_png_is_imwrite_for_a_batch(built):
    """Renderer.iter_png: panels rendered in cv2's BGR order and encoded on the device decode (libpng) to the oracle's
    pixels, like files written by the reference's save_image (src/utils.py:18-21)."""
    import cv2
    from oracle import ref_render
    from reasoning_image_generation_b200 import Renderer
    from reasoning_image_generation_b200.png import PngEncoder
    from reasoning_image_generation_b200.synth import scene_to_elements, synthetic_scene
    sc = synthetic_scene(40, seed=31)
    r = Renderer(512, 512, max_panels=16)
    enc = PngEncoder(512, 512, max_images=16)
    be = ref_render.CBackend()
    seen = 0
    for p0, buf, off in r.iter_png(sc, enc, chunk=16):
        for k in range(len(off) - 1):
            img = cv2.imdecode(np.asarray(buf[int(off[k]):int(off[k + 1])]), cv2.IMREAD_UNCHANGED)      # BGR, as cv2.imread gives
            els, ug = scene_to_elements(sc, p0 + k)
            ref = ref_render.render_panel(512, 512, (255, 255, 255), els, ug, 3, backend=be)
            assert (img == ref).all()
            seen += 1
    assert seen == 40
    r.close()
    enc.close()


def test_two_batches_in_flight(built):
    """rig_png_encode_begin / rig_png_encode_end: begin, begin, end, end gives the same files as two synchronous encodes; a third
    begin without an end is refused; closing a pipelined iterator early leaves the encoder usable."""
    import torch
    from reasoning_image_generation_b200.figure import FigureRenderer
    from reasoning_image_generation_b200.png import PngEncoder, PngError, split_files
    from reasoning_image_generation_b200.synth_figure import synthetic_figure_scene
    sc = synthetic_figure_scene(24, 256, 256, seed=3, kind='multi')
    r = FigureRenderer(256, 256, max_images=32)
    imgs = r.render(sc, 'rgb')
    a = torch.from_numpy(imgs[:12].copy()).cuda()
    b = torch.from_numpy(imgs[12:].copy()).cuda()
    enc = PngEncoder(256, 256, max_images=16)
    ref_a = split_files(*enc.encode(a.data_ptr(), 12))
    ref_b = split_files(*enc.encode(b.data_ptr(), 12))
    for _ in range(3):                                    # odd number of rounds: the slots alternate roles
        sa = enc.encode_begin(a.data_ptr(), 12)
        sb = enc.encode_begin(b.data_ptr(), 12)
        assert {sa, sb} == {0, 1}
        with pytest.raises(PngError):
            enc.encode_begin(a.data_ptr(), 12)
        assert split_files(*enc.encode_end(sa)) == ref_a
        assert split_files(*enc.encode_end(sb)) == ref_b
        sa = enc.encode_begin(a.data_ptr(), 12)           # and a lone begin/end in between
        assert split_files(*enc.encode_end(sa)) == ref_a
    with pytest.raises(PngError):
        enc.encode_end(0)                                 # nothing in flight
    _check(ref_a + ref_b, imgs)
    it = r.iter_png(sc, enc, chunk=4)
    next(it); next(it)
    it.close()                                            # retires the slot whose copy is in flight
    assert split_files(*enc.encode(b.data_ptr(), 12)) == ref_b
    r.close()
    enc.close()


def test_blank_band_hint(built):
    """rig_figure_blank_bands -> rig_png_set_blank_hint: k_figure's flags are truthful (a flagged band is white in the rendered
    pixels), the files encoded with the hint are byte-identical to the files encoded without it, the hint lasts for one call, and
    it is really what the encoder goes by (an all-ones hint gives nearly white files)."""
    import torch
    from reasoning_image_generation_b200.figure import FigureRenderer
    from reasoning_image_generation_b200.png import PngEncoder, split_files
    from reasoning_image_generation_b200.synth_figure import synthetic_figure_scene
    for (W, H) in ((512, 512), (256, 37)):
        n = 40
        sc = synthetic_figure_scene(n, W, H, seed=21, kind='multi')
        r = FigureRenderer(W, H, max_images=64)
        assert r.blank_bands()[0] == 0                       # nothing rendered yet
        r.set_blank_flags(True)
        dev = torch.empty(n * H * W * 3, dtype=torch.uint8, device='cuda')
        r.render_to_device(sc, dev.data_ptr())
        r.sync()
        ptr, rows = r.blank_bands()
        assert ptr and rows >= 1
        bands = (H + rows - 1) // rows
        flags = _dev_bytes(ptr, n * bands).reshape(n, bands)
        imgs = dev.cpu().numpy().reshape(n, H, W, 3)
        assert set(np.unique(flags)) <= {0, 1}
        if H >= 128:
            assert flags.mean() > 0.2                        # the synthetic figures leave blank margins
        for i in range(n):
            for b in np.nonzero(flags[i])[0]:
                assert (imgs[i, b * rows:(b + 1) * rows] == 255).all(), (i, b)
        enc = PngEncoder(W, H, max_images=64, channels=4)
        plain = split_files(*enc.encode(dev.data_ptr(), n))
        enc.set_blank_hint(ptr, rows)
        hinted = split_files(*enc.encode(dev.data_ptr(), n))
        assert hinted == plain
        _check(plain, imgs)
        if H == 512:
            # all-ones hint: once a warp holds its copy of a blank piece (made from pixels it has read: the white images in front
            # give every warp one), it takes the hint's word and the drawn rows of the last images come out white
            from reasoning_image_generation_b200.png import decode_png_strict
            nw = 120
            both = torch.cat([torch.full((nw * H * W * 3,), 255, dtype=torch.uint8, device='cuda'), dev[:8 * H * W * 3]])
            ones = torch.ones((nw + 8) * bands, dtype=torch.uint8, device='cuda')
            enc2 = PngEncoder(W, H, max_images=nw + 8, channels=3)
            enc2.set_blank_hint(ones.data_ptr(), rows)
            files = split_files(*enc2.encode(both.data_ptr(), nw + 8))
            drawn = sum(int((decode_png_strict(f) != 255).any(axis=2).sum()) for f in files[nw:])
            assert drawn < 0.1 * int((imgs[:8] != 255).any(axis=3).sum())
            enc2.close()
        enc.set_blank_hint(ptr, rows)
        enc.set_blank_hint(0, 0)                              # cleared again
        assert split_files(*enc.encode(dev.data_ptr(), n)) == plain      # the hint was for one call
        r.set_blank_flags(False)                              # off again: the next render writes no flags
        r.render_to_device(sc, dev.data_ptr())
        r.sync()
        assert r.blank_bands()[0] == 0
        r.close()
        enc.close()


def _dev_bytes(ptr, nbytes):
    """Device memory at a raw pointer -> numpy (through cudaMemcpy of the runtime torch already loaded)."""
    import ctypes
    import torch
    out = np.zeros(nbytes, np.uint8)
    rt = ctypes.CDLL(None)
    fn = getattr(rt, 'cudaMemcpy', None)
    if fn is None:
        import glob
        import os
        libs = glob.glob(os.path.join(os.path.dirname(torch.__file__), 'lib', 'libcudart*.so*')) + \
            glob.glob(os.path.join(os.path.dirname(torch.__file__), '..', 'nvidia', 'cuda_runtime', 'lib', 'libcudart.so*'))
        fn = ctypes.CDLL(libs[0]).cudaMemcpy
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    fn.restype = ctypes.c_int
    assert fn(out.ctypes.data, ptr, nbytes, 2) == 0
    return out


@pytest.mark.parametrize('rows', [1, 5, 64])
def test_blank_hint_with_other_band_heights(built, rows):
    """The hint's bands need not be the encoder's 2-row pieces (a renderer with 5-row bands, say): a piece is skipped only if
    every band it touches is flagged.  Flags computed from the pixels on the host; files identical to the unhinted ones."""
    import torch
    from reasoning_image_generation_b200.figure import FigureRenderer
    from reasoning_image_generation_b200.png import PngEncoder, split_files
    from reasoning_image_generation_b200.synth_figure import synthetic_figure_scene
    n, W, H = 200, 256, 131
    sc = synthetic_figure_scene(n, W, H, seed=4, kind='single')
    r = FigureRenderer(W, H, max_images=n)
    imgs = r.render(sc, 'rgb')
    r.close()
    bands = (H + rows - 1) // rows
    white_rows = (imgs == 255).all(axis=(2, 3))                              # [n, H]
    flags = np.zeros((n, bands), np.uint8)
    for b in range(bands):
        flags[:, b] = white_rows[:, b * rows:(b + 1) * rows].all(axis=1)
    dev = torch.from_numpy(imgs).cuda()
    fl = torch.from_numpy(flags).cuda()
    enc = PngEncoder(W, H, max_images=n)
    plain = split_files(*enc.encode(dev.data_ptr(), n))
    enc.set_blank_hint(fl.data_ptr(), rows)
    assert split_files(*enc.encode(dev.data_ptr(), n)) == plain
    _check(plain[:20], imgs[:20])
    enc.close()
