"""Host logic of the pipelined PNG loop (png.pipelined_png / png.chunk_plan) with a fake encoder: chunk order, two slots in
flight, early close, errors raised on the worker thread.  No GPU and no library call."""
import threading

import numpy as np
import pytest

from reasoning_image_generation_b200.png import chunk_plan, pipelined_png


class FakeEncoder:
    """Records the calls PngEncoder would get; a 'file' is the bytes of its item number."""

    def __init__(self, chunk):
        self.max_images = chunk
        self.log, self.busy, self.begun = [], {}, 0
        self.hint = None
        self.lock = threading.Lock()

    def staging(self, n):
        self.log.append(('staging', n))
        return 1 << 20

    def wait_event(self, ev):
        self.log.append(('wait', ev))

    def set_blank_hint(self, ptr, rows):
        self.hint = (ptr, rows)

    def encode_begin(self, ptr, n, layout='rgb', stream=0):
        with self.lock:
            slot = self.begun & 1
            assert slot not in self.busy, 'both slots in flight'
            self.busy[slot] = (ptr, n, self.hint)
            self.hint = None
            self.begun += 1
            self.log.append(('begin', slot, ptr, n))
            return slot

    def encode_end(self, slot):
        with self.lock:
            ptr, n, hint = self.busy.pop(slot)
            self.log.append(('end', slot))
        off = np.arange(n + 1, dtype=np.uint64) * 4
        return np.full(4 * n, ptr & 0xff, np.uint8), off

    def encode(self, ptr, n, layout='rgb', stream=0):
        return self.encode_end(self.encode_begin(ptr, n, layout, stream))


def test_chunk_plan_covers_the_batch():
    for n, c in ((0, 8), (5, 8), (8, 8), (9, 8), (32768, 4096), (10000, 4096)):
        plan = chunk_plan(n, c)
        assert [a for a, _ in plan] == list(range(0, n, c))
        assert all(0 < b - a <= c for a, b in plan) and (not plan or plan[-1][1] == n)


@pytest.mark.parametrize('overlap', [True, False])
def test_chunks_come_back_in_order_with_two_slots(overlap):
    enc = FakeEncoder(16)
    rendered = []

    def render_chunk(i0, i1, ptr):
        rendered.append((i0, i1, ptr))
        return i1 - i0, 1000 + i0, (0x5000 + i0, 2)

    got = [(i0, len(off) - 1) for i0, buf, off in pipelined_png(enc, chunk_plan(70, 16), 16, 100, render_chunk, 'rgb', overlap)]
    assert got == [(0, 16), (16, 16), (32, 16), (48, 16), (64, 6)]
    assert [r[:2] for r in rendered] == chunk_plan(70, 16)
    if overlap:
        # the three staging slots take turns, and a chunk is begun before the previous one is ended
        assert [r[2] for r in rendered] == [(1 << 20) + (k % 3) * 16 * 100 for k in range(5)]
        kinds = [e[0] for e in enc.log if e[0] in ('begin', 'end')]
        assert kinds == ['begin', 'begin', 'end', 'begin', 'end', 'begin', 'end', 'begin', 'end', 'end']
    assert not enc.busy
    assert [e[1] for e in enc.log if e[0] == 'wait'] == [1000, 1016, 1032, 1048, 1064]


def test_closing_early_retires_the_slot_and_stops_the_worker():
    enc = FakeEncoder(8)
    calls = []

    def render_chunk(i0, i1, ptr):
        calls.append(i0)
        return i1 - i0, 0

    it = pipelined_png(enc, chunk_plan(80, 8), 8, 10, render_chunk)
    next(it)
    next(it)
    it.close()
    assert not enc.busy                       # the chunk whose copy was in flight has been ended
    assert len(calls) <= 6                    # the worker ran at most three chunks ahead, then stopped


def test_worker_errors_surface_on_the_consumer():
    enc = FakeEncoder(8)

    def render_chunk(i0, i1, ptr):
        if i0 == 16:
            raise RuntimeError('render failed')
        return i1 - i0, 0

    seen = []
    with pytest.raises(RuntimeError, match='render failed'):
        for i0, buf, off in pipelined_png(enc, chunk_plan(40, 8), 8, 10, render_chunk):
            seen.append(i0)
    assert seen == [0] and not enc.busy       # chunk 8 was begun but its files never handed out; its slot is retired
