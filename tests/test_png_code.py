"""The Huffman code of the device-side PNG writer (csrc/rig_png_code.h), checked on the CPU: the block header the library
builds and a token stream coded with its tables must inflate (zlib) to the bytes the tokens stand for -- the pin of everything
in the file format that does not depend on the GPU."""
import ctypes
import zlib

import numpy as np
import pytest

from reasoning_image_generation_b200 import _abi


def code_tables(ch):
    lib = _abi.load()
    lib.rig_png_debug_code.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.rig_png_debug_code.restype = ctypes.c_int
    lens = np.zeros(286, np.uint8)
    hdr = np.zeros(96, np.uint32)
    hbits = ctypes.c_uint32(0)
    lut = np.zeros(516, np.uint32)
    assert lib.rig_png_debug_code(ch, lens.ctypes.data, hdr.ctypes.data, ctypes.byref(hbits), lut.ctypes.data) == 0
    return lens, hdr, int(hbits.value), lut


class Bits:
    def __init__(self, words, nbits):
        self.acc = 0
        for i, w in enumerate(words):
            self.acc |= int(w) << (32 * i)
        self.acc &= (1 << nbits) - 1
        self.n = nbits

    def put(self, entry):
        entry = int(entry)
        self.acc |= (entry & 0xffffff) << self.n
        self.n += entry >> 24

    def bytes(self):
        return self.acc.to_bytes((self.n + 7) // 8, 'little')


def tokens(row, lut, b):
    """The writer's token rule for one filtered row (serial definition, tools/png_parallel_walk_proto.py)."""
    i, n = 0, len(row)
    while i < n:
        if row[i]:
            b.put(lut[int(row[i])]); i += 1; continue
        j = i
        while j < n and row[j] == 0:
            j += 1
        r = j - i - 1
        b.put(lut[0])
        while r >= 258:
            b.put(lut[256 + 258]); r -= 258
        if r >= 3:
            b.put(lut[256 + r]); r = 0
        while r:
            b.put(lut[0]); r -= 1
        i = j


@pytest.mark.parametrize('ch', [3, 4])
def test_code_is_a_complete_prefix_code(built, ch):
    lens, hdr, hbits, lut = code_tables(ch)
    assert lens.min() >= 1 and lens.max() <= 15
    assert sum(2.0 ** -int(l) for l in lens) == 1.0                 # Kraft: complete, as inflate demands of the literal/length code
    # the frequent tokens are the short ones
    assert lens[285] <= 3 and lens[0] <= 4 and lens[255] <= 4 and lens[1] <= 5
    assert 0 < hbits <= 96 * 32
    # every table entry fits its 24 bits
    assert all((int(e) >> 24) <= 24 and (int(e) & 0xffffff) < (1 << (int(e) >> 24)) for e in lut if e)


@pytest.mark.parametrize('ch', [3, 4])
def test_header_plus_tokens_inflate(built, ch):
    lens, hdr, hbits, lut = code_tables(ch)
    rng = np.random.RandomState(ch)
    rows = []
    for k in range(40):
        n = 1 + ch * 512
        row = np.zeros(n, np.uint8)
        row[0] = 1
        row[1:1 + ch] = 255
        for _ in range(int(rng.randint(0, 9))):
            p = int(rng.randint(1, n))
            row[p:p + int(rng.randint(1, 9))] = rng.randint(0, 256)
        if k == 7:
            row[1:] = rng.randint(0, 256, n - 1)                    # an incompressible row: every literal, long codes
        if k == 9:
            row[1:] = 0
        rows.append(row)
    rows.append(np.r_[1, np.zeros(258 * 3 + 2, np.uint8)])         # run lengths around the 258 boundary
    rows.append(np.r_[1, np.zeros(258 * 3 + 3, np.uint8), 7])
    rows.append(np.r_[1, np.zeros(259, np.uint8), 9, 0, 0, 3, 0, 5])
    b = Bits(hdr, hbits)
    for r in rows:
        tokens(r, lut, b)
    b.put(lut[515])                                                 # end of block
    raw = zlib.decompress(b.bytes(), wbits=-15)
    assert raw == b''.join(bytes(r) for r in rows)


def test_every_match_length_decodes(built):
    lens, hdr, hbits, lut = code_tables(4)
    b = Bits(hdr, hbits)
    want = b''
    for l in range(3, 259):
        b.put(lut[77]); b.put(lut[0]); b.put(lut[256 + l])
        want += bytes([77]) + bytes(1 + l)
    b.put(lut[515])
    assert zlib.decompress(b.bytes(), wbits=-15) == want
