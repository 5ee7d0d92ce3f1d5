"""Prototype (CPU, numpy) of a byte-parallel formulation of the encoder's token walk, checked against the serial definition
k_png_pieces implements and against zlib.

Serial definition (one filtered row F, whose first byte is the filter type, never zero): a non-zero byte is a literal; a maximal
run of r zeros is a literal zero followed by png_zero_matches(r - 1): matches of 258 at distance 1 while >= 258 remain, one match of
the remainder if it is >= 3, else one or two literal zeros.

Parallel formulation: every byte knows the start s and the length r of its zero run (segmented scans on the device); with
k = i - s, q = r - 1, the byte starts a token iff it is non-zero, or k == 0, or (k - 1) % 258 == 0 and k - 1 < 258 * (q // 258),
or k == 1 + 258 * (q // 258) with rest = q % 258 >= 3, or it lies in the rest and rest < 3.  bits(i) = the length of that token,
else 0; the bit offset is the exclusive prefix sum; every token-starting byte writes its own code.  No lane walks a chunk.
"""
import zlib

import numpy as np


def rev(code, n):
    return int(format(code, '0%db' % n)[::-1], 2)


def literal(v):
    return (rev(0x30 + v, 8), 8) if v < 144 else (rev(0x190 + v - 144, 9), 9)


def match(length):                      # length 3..258 at distance 1: symbol + extra bits + 5-bit distance code 0
    if length == 258:
        sym, eb, extra = 285, 0, 0
    else:
        l = length - 3
        if l < 8:
            sym, eb, extra = 257 + l, 0, 0
        else:
            eb = l.bit_length() - 1 - 2
            sym, extra = 257 + 4 * eb + 4 + ((l >> eb) & 3), l & ((1 << eb) - 1)
    code, n = (rev(sym - 256, 7), 7) if sym < 280 else (rev(0xc0 + sym - 280, 8), 8)
    return code | (extra << n), n + eb + 5


class Bits:
    def __init__(self):
        self.acc, self.n = 0, 0

    def put(self, code, n):
        self.acc |= code << self.n
        self.n += n


def serial(F):
    b = Bits()
    i, n = 0, len(F)
    while i < n:
        if F[i]:
            b.put(*literal(int(F[i])))
            i += 1
            continue
        j = i
        while j < n and F[j] == 0:
            j += 1
        r = j - i - 1
        b.put(*literal(0))
        while r >= 258:
            b.put(*match(258))
            r -= 258
        if r >= 3:
            b.put(*match(r))
            r = 0
        while r:
            b.put(*literal(0))
            r -= 1
        i = j
    return b


def parallel(F):
    F = np.asarray(F, np.int64)
    n = len(F)
    idx = np.arange(n)
    nzpos = np.where(F != 0, idx, -1)
    last_nz = np.maximum.accumulate(nzpos)                       # last non-zero at or before i
    nxt = np.where(F != 0, idx, n)
    next_nz = np.minimum.accumulate(nxt[::-1])[::-1]             # first non-zero at or after i
    s = last_nz + 1                                              # run start (for zero bytes)
    r = next_nz - s                                              # run length
    k = idx - s
    q = r - 1
    n258 = q // 258
    rest = q % 258
    codes = np.zeros(n, np.int64)
    bits = np.zeros(n, np.int64)
    for i in range(n):                                           # (element-wise; a device thread per byte)
        if F[i]:
            codes[i], bits[i] = literal(int(F[i]))
        elif k[i] == 0:
            codes[i], bits[i] = literal(0)
        elif (k[i] - 1) % 258 == 0 and k[i] - 1 < 258 * n258[i]:
            codes[i], bits[i] = match(258)
        elif k[i] == 1 + 258 * n258[i] and rest[i] >= 3:
            codes[i], bits[i] = match(int(rest[i]))
        elif k[i] >= 1 + 258 * n258[i] and rest[i] < 3:
            codes[i], bits[i] = literal(0)
    off = np.cumsum(bits) - bits
    b = Bits()
    for i in range(n):
        if bits[i]:
            b.acc |= int(codes[i]) << int(off[i])
    b.n = int(bits.sum())
    return b


def as_deflate(b):                      # BFINAL=1, BTYPE=01 + tokens + end of block
    acc = 3 | (b.acc << 3)
    n = b.n + 3 + 7
    return acc.to_bytes((n + 7) // 8, 'little')


rng = np.random.RandomState(1)
cases = []
for _ in range(300):
    n = int(rng.choice([1, 2, 5, 49, 258, 259, 260, 261, 262, 517, 518, 520, 1537, 2049]))
    F = np.zeros(n, np.int64)
    F[0] = 1
    for _ in range(int(rng.randint(0, 12))):
        p = int(rng.randint(0, n))
        F[p:p + int(rng.randint(1, 6))] = rng.randint(0, 256)
    F[0] = 1
    cases.append(F)
cases.append(np.r_[1, rng.randint(0, 256, 300)])
cases.append(np.r_[1, np.zeros(2048, np.int64)])
for F in cases:
    a, p = serial(F), parallel(F)
    assert (a.acc, a.n) == (p.acc, p.n), F
    raw = zlib.decompress(as_deflate(p), wbits=-15)
    assert raw == bytes(int(v) for v in F)
print('serial == parallel on %d rows; all decode with zlib' % len(cases))
