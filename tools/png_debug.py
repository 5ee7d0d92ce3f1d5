"""Debug aid: encode small images, inflate the raw deflate stream (no Adler check) and compare with the expected filtered bytes."""
import os, sys, zlib, struct
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from reasoning_image_generation_b200.png import PngEncoder, split_files

def expected(img, ch):
    H, W, _ = img.shape
    a = img.astype(np.int16)
    if ch == 4:
        a = np.concatenate([a, np.full((H, W, 1), 255, np.int16)], axis=2)
    rows = a.reshape(H, W * ch)
    left = np.concatenate([np.zeros((H, ch), np.int16), rows[:, :-ch]], axis=1)
    F = ((rows - left) & 255).astype(np.uint8)
    return b''.join(bytes([1]) + bytes(F[y]) for y in range(H))

def run(imgs, ch):
    n, H, W, _ = imgs.shape
    t = torch.from_numpy(np.ascontiguousarray(imgs)).cuda()
    enc = PngEncoder(W, H, max_images=n, channels=ch)
    buf, off = enc.encode(t.data_ptr(), n, 'rgb')
    files = split_files(buf, off)
    enc.close()
    for i, f in enumerate(files):
        pos, idat = 8, b''
        while pos < len(f):
            (ln,), typ = struct.unpack('>I', f[pos:pos + 4]), f[pos + 4:pos + 8]
            if typ == b'IDAT': idat += f[pos + 8:pos + 8 + ln]
            pos += 12 + ln
        d = zlib.decompressobj(-15)
        try:
            raw = d.decompress(idat[2:])
        except Exception as e:
            print(f'ch={ch} img {i}: inflate error {e}; file {len(f)} B'); continue
        exp = expected(imgs[i], ch)
        ok = raw == exp
        adler_ok = struct.unpack('>I', idat[-4:])[0] == (zlib.adler32(exp) & 0xffffffff)
        msg = ''
        if not ok:
            k = next((j for j in range(min(len(raw), len(exp))) if raw[j] != exp[j]), min(len(raw), len(exp)))
            RL = 1 + ch * W
            msg = f' first mismatch at byte {k} (row {k // RL}, col {k % RL}); len raw {len(raw)} exp {len(exp)}'
        print(f'ch={ch} img {i}: {len(f)} B, content {"OK" if ok else "BAD"}, adler {"OK" if adler_ok else "BAD"}, unused {len(d.unused_data)}{msg}')

rng = np.random.RandomState(0)
W, H = 512, 8
imgs = np.full((6, H, W, 3), 255, np.uint8)
imgs[1, 3, 100] = (10, 20, 30)
imgs[2, :, 200:210] = 0
imgs[3] = rng.randint(0, 256, (H, W, 3))
imgs[4, 2] = rng.randint(0, 256, (W, 3))
imgs[5, 5, 300:340] = rng.randint(0, 256, (40, 3))
for ch in (4, 3):
    run(imgs, ch)
