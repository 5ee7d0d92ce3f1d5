"""Token statistics of the PNG writer's stream over synthetic figures (CPU, uses the oracle renderer: test tooling, not product):
the frequency model in csrc/rig_png_code.h was read off this table.  Also prints the file size a Huffman code limited to
L bits would give, for L = 15, 12, 11, 10, 9.    python tools/png_token_stats.py [n_images]"""
import collections
import heapq
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.agg_oracle import AggOracle, build_agg_oracle  # noqa: E402
from reasoning_image_generation_b200.synth_figure import synthetic_figure_scene  # noqa: E402


def match_sym(length):
    length = int(length)
    if length == 258:
        return 285, 0
    l = length - 3
    if l < 8:
        return 257 + l, 0
    eb = l.bit_length() - 1 - 2
    return 257 + 4 * eb + 4 + ((l >> eb) & 3), eb


def tokens_row(F, hist, extra):
    nz = np.flatnonzero(F)
    prev = -1
    for i in list(nz) + [len(F)]:
        r = i - prev - 1
        if r > 0:
            hist[0] += 1
            r -= 1
            hist[285] += r // 258
            hist['d'] += r // 258
            r %= 258
            if r >= 3:
                s, eb = match_sym(r)
                hist[s] += 1
                hist['d'] += 1
                extra[0] += eb
            else:
                hist[0] += r
        if i < len(F):
            hist[int(F[i])] += 1
        prev = i


def limited_lengths(freqs, maxbits):
    syms = list(freqs)
    h = [(f, i, (s,)) for i, (s, f) in enumerate(freqs.items())]
    heapq.heapify(h)
    L = {s: 0 for s in syms}
    c = len(h)
    while len(h) > 1:
        a, b = heapq.heappop(h), heapq.heappop(h)
        for s in a[2] + b[2]:
            L[s] += 1
        heapq.heappush(h, (a[0] + b[0], c, a[2] + b[2]))
        c += 1
    full = 1 << maxbits
    for s in syms:
        L[s] = min(L[s], maxbits)
    k = sum(full >> L[s] for s in syms)
    order = sorted(syms, key=lambda s: freqs[s])
    while k > full:
        for s in order:
            if L[s] < maxbits:
                k -= full >> (L[s] + 1)
                L[s] += 1
                break
    while k < full:
        for s in reversed(order):
            if L[s] > 1 and k + (full >> L[s]) <= full:
                k += full >> L[s]
                L[s] -= 1
                break
        else:
            break
    return L


if __name__ == '__main__':
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    build_agg_oracle()
    o = AggOracle()
    for kind in ('single', 'multi'):
        sc = synthetic_figure_scene(N, 512, 512, seed=1000, kind=kind)
        hist, extra = collections.Counter(), [0]
        for k in range(N):
            a = np.asarray(o.render_rgba(sc, k, 512, 512))
            if a.shape[2] == 3:
                a = np.concatenate([a, np.full(a.shape[:2] + (1,), 255, np.uint8)], axis=2)
            rows = a.reshape(512, 2048).astype(np.int16)
            left = np.concatenate([np.zeros((512, 4), np.int16), rows[:, :-4]], axis=1)
            F = ((rows - left) & 255).astype(np.uint8)
            for y in range(512):
                tokens_row(np.concatenate([[1], F[y]]), hist, extra)
        syms = {s: hist.get(s, 0) + 1 for s in range(286)}
        print(kind, 'tokens/image', sum(v for k, v in hist.items() if k != 'd') / N)
        for lim in (15, 12, 11, 10, 9):
            L = limited_lengths(syms, lim)
            bits = sum(hist.get(s, 0) * L[s] for s in range(286)) + hist['d'] + extra[0]
            print('  limit %2d bits: %.2f KB/image; lengths of 285, 0, 255, 1, 284: %s' % (lim, bits / N / 8 / 1e3, [L[s] for s in (285, 0, 255, 1, 284)]))
        print('  top symbols', sorted(((s, hist[s]) for s in hist if s != 'd'), key=lambda kv: -kv[1])[:14])
