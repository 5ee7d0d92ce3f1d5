"""driver.py -- batched drop-in for the reference's dataset runner (src/cli.py:13-34,82-103 + src/generator.py:210-314).

The reference renders, PNG-encodes and composes one panel at a time inside every pool worker.  Here the two halves are split
the way SURVEY 8(b) recommends:

* **emit** (host, pool workers, never touches CUDA): the reference's own `RPMGenerator` runs unchanged -- same rule functions,
  same seeds, same `random` stream -- but the three pixel seams are replaced by recorders:
    `RPMGenerator._render_elements_to_canvas` (generator.py:66)  -> appends a panel to a `SceneBuilder` (the rand_color() draws
                                                                     for colourless elements happen here, at call time) and
                                                                     returns a `PanelRef` instead of pixels
    `save_image` (utils.py:18-21, bound in generator.py:17)       -> records (path, panel)
    `compose_grid` (layout.py:49-192, bound in generator.py:21)   -> records which panels the overview shows and returns the
                                                                     reference's metadata without drawing
  The rule logic never looks at pixels (generator.py:338-541 only stores, saves and composes them), so the emitted element
  lists, labels, meta.json / coco.json are those of the reference bit for bit.
* **own** (one process per GPU): concatenates the emitted scenes, renders them with librig (k_expand / k_raster), composes the
  overviews (k_grid_*), encodes PNG files on the device (k_png_*) and writes them under the reference's layout
  `samples/sample_%06d/{state_*,proto_true_next,option_*,query}.png`, `grids/grid_%06d.png`, `index.json`.

B2: `rule_rotate` / `rule_flip` (rules.py:223-228,360-365) spin forever when the only selectable element
is a circle / sits in the centre cell; the reference loses such a sample after its 30 s watchdog (generator.py:180-207) and
leaves the thread spinning.  `_would_spin` evaluates the loops' exit condition up front: when no choice of `idx` can ever leave
the loop the sample is failed at once with the reference's own TimeoutException (same outcome: `None`, directory cleaned), and
otherwise the reference loop runs unchanged (so the draws it consumes are unchanged).
"""
from __future__ import annotations

import importlib
import json
import os
import sys
import threading
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .scene import Scene, SceneBuilder


class PanelRef:
    """What the emit stage hands the reference in place of a rendered panel: the index of the panel in the sample's scene."""
    __slots__ = ('index',)

    def __init__(self, index: int):
        self.index = int(index)


class GridRef:
    """Stands for the sample's overview image (compose_grid's first return value)."""
    __slots__ = ('states', 'options')

    def __init__(self, states: List[int], options: List[int]):
        self.states, self.options = states, options


class _Spin(BaseException):
    """Raised by the B2 guard.  A BaseException so that the reference's `except Exception` fallbacks around the rule handlers
    (generator.py:372-385,476-484) do not swallow it -- in the reference nothing interrupts the loop either."""


class Tape:
    """Everything one sample emitted."""

    def __init__(self, W: int, H: int, bg_color, grid_size: int):
        self.builder = SceneBuilder(W, H, bg_color, grid_size)
        self.files: List[Tuple[str, int]] = []          # (path, panel index)
        self.grid: Optional[GridRef] = None
        self.grid_path: Optional[str] = None
        self.query_path: Optional[str] = None
        self.layout_key: Optional[Tuple] = None         # arguments of the overview layout (GridLayout(*key))
        self.panels: List[Tuple[list, bool]] = []       # (renderer view of the elements, use_grid), for fixtures / tests


_local = threading.local()

# the fields _render_elements_to_canvas reads (generator.py:87-94)
RENDER_FIELDS = ('kind', 'size', 'fill', 'stroke_width', 'center', 'color', 'angle')


def _render_view(el: Dict[str, Any]) -> Dict[str, Any]:
    out = {}
    for k in RENDER_FIELDS:
        if k in el:
            v = el[k]
            if isinstance(v, (tuple, list)):
                v = [x.item() if isinstance(x, np.generic) else x for x in v]
            elif isinstance(v, np.generic):
                v = v.item()
            out[k] = v
    return out


def _would_spin(leaf_fn_name: str, history_elements, rule_info, config, grid_size: int = 3) -> bool:
    elems = history_elements[-1] if history_elements else []
    if not elems:
        return False                                     # rule_rotate returns early; rule_flip raises IndexError (caught upstream)
    info = rule_info if rule_info else {}
    cand = [info['idx']] if 'idx' in info else range(len(elems))
    if leaf_fn_name == 'rule_rotate':
        return all(elems[i].get('kind', None) == 'circle' for i in cand)
    W, H = config.canvas_size
    cw, ch = float(W) / grid_size, float(H) / grid_size
    mid = grid_size // 2
    return all(elems[i]['center'][0] // cw == mid and elems[i]['center'][1] // ch == mid for i in cand)


def _guard(fn):
    name = fn.__name__

    def guarded(history_elements, rule_info=None, config=None, **kw):
        try:
            spin = config is not None and _would_spin(name, history_elements, rule_info, config, kw.get('grid_size', 3))
        except (IndexError, KeyError, TypeError):
            spin = False                                 # malformed input: let the reference raise what it raises
        if spin:
            raise _Spin(f'{name}: no selectable element can leave the selection loop (rules.py B2)')
        return fn(history_elements, rule_info=rule_info, config=config, **kw)
    guarded.__name__ = name                              # handler_name in meta.json is fn.__name__ (generator.py:255)
    guarded.__wrapped__ = fn
    return guarded


def _compose_meta(W, H, states, candidates, sample_dir, num_options, margin=20, padding_v=20, show_labels=True, show_border=True,
                  bg_color=(255, 255, 255)):
    """compose_grid's return values (layout.py:66-192) without the pixels."""
    cols_seq, cols_opt = len(states) + 1, num_options
    max_cell_w = (W - 2 * margin) // max(1, max(cols_seq, cols_opt))
    max_cell_h = (H - 2 * margin - padding_v) // 2
    cell = max(1, min(max_cell_w, max_cell_h))
    grid_h = 2 * cell + padding_v + 2 * margin
    seq_off, opt_off = (W - cols_seq * cell) // 2, (W - cols_opt * cell) // 2
    top_y, bottom_y = margin, margin + cell + padding_v
    query_path = os.path.join(sample_dir, 'query.png')
    cells_meta = []
    for i in range(cols_seq):
        q = i == len(states)
        cells_meta.append({'r': 0, 'c': i, 'label': f'S{i}' if show_labels else '', 'bbox': [int(seq_off + i * cell), int(top_y), int(cell), int(cell)],
                           'proto_path': None if q else states[i].get('proto_path'), 'state_path': None if q else states[i].get('state_path'),
                           'is_query': bool(q), 'query_path': query_path if q else None})
    for i in range(cols_opt):
        c = candidates[i]                                # IndexError when fewer candidates than options, as in layout.py:159
        cells_meta.append({'r': 1, 'c': i, 'label': chr(65 + i) if show_labels else '', 'bbox': [int(opt_off + i * cell), int(bottom_y), int(cell), int(cell)],
                           'path': c.get('path'), 'is_correct': bool(c.get('is_correct', False))})
    seq_meta = [{'proto_path': s.get('proto_path'), 'state_path': s.get('state_path'), 'is_query': False} for s in states]
    seq_meta.append({'proto_path': None, 'state_path': None, 'is_query': True, 'query_path': query_path})
    opts_meta = [{'path': c.get('path'), 'is_correct': bool(c.get('is_correct', False)), 'label': chr(65 + k)} for k, c in enumerate(candidates)]
    return cells_meta, seq_meta, opts_meta, query_path, grid_h, cell


def install_emitter(refgen, refrules=None):
    """Patches the reference's `generator` module (and `rules.RULE_MAP`) in this process.  Idempotent."""
    if getattr(refgen, '_rig_emitter', False):
        return refgen
    cls = refgen.RPMGenerator
    impl = cls._generate_sample_impl

    def _panel_of(tape: Tape, img) -> int:
        if isinstance(img, PanelRef):
            return img.index
        # the only pixels the reference makes itself: the all-white fallback of a failed distractor render (generator.py:494)
        a = np.asarray(img)
        if a.dtype == np.uint8 and a.ndim == 3 and (a == 255).all():
            tape.panels.append(([], False))
            return tape.builder.add_panel([], False)
        raise TypeError('emit stage: an image that was not produced by the render seam')

    def render(self, W, H, bg_color, elements, use_grid=False, grid_size=3):
        tape: Tape = _local.tape
        if (W, H) != (tape.builder.W, tape.builder.H):
            raise ValueError('canvas size changed within a sample')
        idx = tape.builder.add_panel(elements, use_grid)   # rand_color() draws happen here (B8)
        tape.panels.append(([_render_view(e) for e in elements], bool(use_grid)))
        return PanelRef(idx)

    def save_image(im, path):
        tape: Tape = _local.tape
        refgen.ensure_dir(os.path.dirname(path))
        if isinstance(im, GridRef):
            tape.grid_path = path
        else:
            tape.files.append((path, _panel_of(tape, im)))

    def compose_grid(W, H, states, candidates, sample_dir, num_options, margin=20, padding_v=20, show_labels=True, show_border=True,
                     bg_color=(255, 255, 255)):
        tape: Tape = _local.tape
        cells_meta, seq_meta, opts_meta, query_path, grid_h, cell = _compose_meta(W, H, states, candidates, sample_dir, num_options, margin,
                                                                                    padding_v, show_labels, show_border, bg_color)
        tape.grid = GridRef([_panel_of(tape, s['state_img']) for s in states], [_panel_of(tape, candidates[i]['img']) for i in range(num_options)])
        tape.query_path = query_path
        tape.layout_key = (W, H, len(states), num_options, margin, padding_v, bool(show_labels), bool(show_border), tuple(bg_color))
        return tape.grid, cells_meta, seq_meta, opts_meta, query_path, grid_h, cell

    def generate_sample_impl(self, sample_id, *a, **kw):
        W, H = self.cfg.canvas_size
        self._rig_tape = _local.tape = Tape(W, H, self.cfg.bg_color, getattr(self.cfg, 'grid_size', 3))
        try:
            return impl(self, sample_id, *a, **kw)
        except _Spin as e:
            raise refgen.TimeoutException(str(e))        # what the reference's watchdog reports 30 s later

    cls._render_elements_to_canvas = render
    cls._generate_sample_impl = generate_sample_impl
    refgen.save_image = save_image
    refgen.compose_grid = compose_grid
    rules = refrules if refrules is not None else refgen.rules
    for leaf, fn in list(rules.RULE_MAP.items()):
        if getattr(fn, '__name__', '') in ('rule_rotate', 'rule_flip') and not hasattr(fn, '__wrapped__'):
            rules.RULE_MAP[leaf] = _guard(fn)
    refgen._rig_emitter = True
    return refgen


def load_reference(ref_src: str):
    """Imports the reference's `generator` / `config` modules from `ref_src` (its src/ directory) and installs the emitter."""
    ref_src = os.path.abspath(ref_src)
    if ref_src not in sys.path:
        sys.path.insert(0, ref_src)
    refgen = importlib.import_module('generator')
    refcfg = importlib.import_module('config')
    if os.path.dirname(os.path.abspath(refgen.__file__)) != ref_src:
        raise ImportError(f'`generator` resolved to {refgen.__file__}, not to {ref_src}')
    install_emitter(refgen)
    return refgen, refcfg


class Emission:
    """Picklable result of one sample: scene arrays + file plan + the reference's meta."""

    def __init__(self, index: int, meta: Optional[Dict[str, Any]], tape: Optional[Tape]):
        self.index = int(index)
        self.meta = meta
        self.ok = meta is not None and not meta.get('error') and tape is not None and tape.grid is not None
        if self.ok:
            self.scene: Optional[Scene] = tape.builder.finish()
            self.files = list(tape.files)
            self.grid_states, self.grid_options = list(tape.grid.states), list(tape.grid.options)
            self.grid_path, self.query_path, self.layout_key = tape.grid_path, tape.query_path, tape.layout_key
            self.panels = tape.panels
        else:
            self.scene, self.files, self.grid_states, self.grid_options = None, [], [], []
            self.grid_path = self.query_path = self.layout_key = None
            self.panels = []


_ref_src_for_workers: Optional[str] = None


def _worker_init(ref_src: str, quiet: bool):
    global _ref_src_for_workers
    _ref_src_for_workers = ref_src
    if quiet:
        import logging
        logging.disable(logging.CRITICAL)
    load_reference(ref_src)


def emit_sample(index: int, out_dir: str, grid: int, seed: Optional[int]) -> Emission:
    """The pool task: cli._make_sample (cli.py:13-34) with the pixel seams recording instead of drawing."""
    refgen, refcfg = load_reference(_ref_src_for_workers)
    try:
        local_seed = None if seed is None else seed + index
        cfg = refcfg.GenConfig(out_dir=out_dir, grid_size=grid, seed=local_seed)
        gen = refgen.RPMGenerator(cfg)
        meta = gen.generate_sample(index)
        return Emission(index, meta, getattr(gen, '_rig_tape', None))
    except Exception as e:                               # cli.py:25-34: an error record instead of a meta
        import traceback
        return Emission(index, {'index': index, 'error': True, 'error_type': str(type(e)), 'error_message': str(e),
                                'traceback': traceback.format_exc()}, None)


def emit_dataset(ref_src: str, out_dir: str, n: int, seed: Optional[int], grid: int = 3, workers: Optional[int] = None,
                 quiet: bool = True, chunksize: int = 8):
    """Yields the `Emission` of samples 0..n-1 in index order.  workers > 1: a process pool created with the 'fork' start method
    BEFORE any CUDA work of the caller, or 'spawn' when this process already holds a CUDA context (fork + CUDA is undefined)."""
    os.makedirs(out_dir, exist_ok=True)
    workers = workers if workers and workers > 0 else (os.cpu_count() or 1)
    if workers == 1:
        _worker_init(ref_src, quiet)
        for i in range(n):
            yield emit_sample(i, out_dir, grid, seed)
        return
    import multiprocessing as mp
    method = 'fork'
    if 'torch' in sys.modules:
        import torch
        if torch.cuda.is_initialized():
            method = 'spawn'
    ctx = mp.get_context(method)
    with ctx.Pool(workers, initializer=_worker_init, initargs=(ref_src, quiet)) as pool:
        yield from pool.imap(_emit_star, [(i, out_dir, grid, seed) for i in range(n)], chunksize=chunksize)


def _emit_star(args):
    return emit_sample(*args)


# ------------------------------------------------------------------------------------------------------------------------
# owner stage: one process per GPU


class DatasetWriter:
    """Renders, composes, encodes and writes the samples of `Emission`s in batches on one GPU."""

    def __init__(self, W: int = 512, H: int = 512, device: int = 0, batch_panels: int = 4096, write_files: bool = True):
        from .png import PngEncoder
        from .render import Renderer
        self.W, self.H, self.device = int(W), int(H), int(device)
        self.batch_panels = int(batch_panels)
        self.write = bool(write_files)
        self.renderer = Renderer(W, H, device=device, max_panels=self.batch_panels, max_elements=8 * self.batch_panels)
        self.panel_png = PngEncoder(W, H, device=device, max_images=self.batch_panels, channels=3)
        self._grid: Dict[Tuple, Tuple] = {}              # layout key -> (GridLayout, GridCompositor, PngEncoder, query.png bytes)
        self._pending: List[Emission] = []
        self._pending_panels = 0
        self.samples = self.panels = self.files = self.png_bytes = 0

    def close(self):
        self.flush()
        for _, comp, enc, _ in self._grid.values():
            comp.close()
            enc.close()
        self._grid.clear()
        self.panel_png.close()
        self.renderer.close()

    def _grid_for(self, key):
        g = self._grid.get(key)
        if g is None:
            import cv2
            from .compose import GridCompositor, GridLayout, _make_query_image
            from .png import PngEncoder
            W, H, n_states, num_options, margin, padding_v, show_labels, show_border, bg = key
            layout = GridLayout(W, H, n_states=n_states, num_options=num_options, margin=margin, padding_v=padding_v,
                                show_labels=show_labels, show_border=show_border, bg_color=bg, panel_size=W)
            max_samples = max(1, self.batch_panels // max(1, n_states + num_options)) + 1
            comp = GridCompositor(layout, device=self.device, max_samples=max_samples, channel_order='bgr')
            enc = PngEncoder(layout.grid_w, layout.grid_h, device=self.device, max_images=max_samples, channels=3)
            # query.png is a constant of the layout (layout.py:33-47,119): the reference's own cv2 calls, encoded once
            ok, q = cv2.imencode('.png', _make_query_image(layout.cell_size))
            g = self._grid[key] = (layout, comp, enc, bytes(q) if ok else b'')
        return g

    def add(self, em: Emission):
        if not em.ok:
            return
        n = em.scene.n_panels
        if self._pending and self._pending_panels + n > self.batch_panels:
            self.flush()
        self._pending.append(em)
        self._pending_panels += n

    def flush(self):
        ems, self._pending, self._pending_panels = self._pending, [], 0
        if not ems:
            return
        import torch
        from .png import write_files
        base, p = [], 0
        for em in ems:
            base.append(p)
            p += em.scene.n_panels
        scene = concat_scenes([em.scene for em in ems])
        with torch.cuda.device(self.device):
            dev = torch.empty((scene.n_panels, self.H, self.W, 3), dtype=torch.uint8, device=f'cuda:{self.device}')
            self.renderer.render_to_device_checked(scene, dev.data_ptr(), layout='bgr')
            buf, off = self.panel_png.encode(dev.data_ptr(), scene.n_panels, layout='bgr')
            paths, which = [], []
            for em, b in zip(ems, base):
                for path, k in em.files:
                    paths.append(path)
                    which.append(b + k)
            if self.write:
                for path, k in zip(paths, which):
                    with open(path, 'wb') as f:
                        f.write(memoryview(buf[int(off[k]):int(off[k + 1])]))
            self.files += len(paths)
            self.png_bytes += int(sum(int(off[k + 1]) - int(off[k]) for k in which))
            by_key: Dict[Tuple, List[int]] = {}
            for s, em in enumerate(ems):
                by_key.setdefault(em.layout_key, []).append(s)
            for key, members in by_key.items():
                layout, comp, enc, qbytes = self._grid_for(key)
                for c0 in range(0, len(members), comp.max_samples):
                    part = members[c0:c0 + comp.max_samples]
                    idx = np.array([[base[s] + k for k in ems[s].grid_states + ems[s].grid_options] for s in part], dtype=np.int32)
                    gdev = torch.empty((len(part), layout.grid_h, layout.grid_w, 3), dtype=torch.uint8, device=f'cuda:{self.device}')
                    comp.compose_to_device(dev.data_ptr(), idx, gdev.data_ptr())
                    comp.sync()
                    gbuf, goff = enc.encode(gdev.data_ptr(), len(part), layout='bgr')
                    if self.write:
                        write_files(gbuf, goff, [ems[s].grid_path for s in part])
                        for s in part:
                            with open(ems[s].query_path, 'wb') as f:
                                f.write(qbytes)
                    self.files += 2 * len(part)
                    self.png_bytes += int(goff[len(part)])
        self.samples += len(ems)
        self.panels += scene.n_panels


def concat_scenes(scenes: Sequence[Scene]) -> Scene:
    first = scenes[0]
    begin = [np.zeros(1, np.uint32)]
    e = 0
    for s in scenes:
        if (s.W, s.H, s.bg_color, s.grid_size) != (first.W, first.H, first.bg_color, first.grid_size):
            raise ValueError('scenes of one batch must share canvas, background and grid size')
        begin.append(s.begin[1:] + np.uint32(e))
        e += s.n_elements

    def cat(name):
        return np.concatenate([getattr(s, name) for s in scenes])
    return Scene(first.W, first.H, first.bg_color, first.grid_size, np.concatenate(begin), cat('flags'), cat('kind'), cat('fill'), cat('flip'),
                 cat('stroke'), cat('size'), cat('angle'), cat('cx'), cat('cy'), cat('rgb'))


def write_index(out_dir: str, metas):
    """cli.write_index (cli.py:49-52)."""
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, 'index.json'), 'w', encoding='utf-8') as f:
        json.dump(metas, f, ensure_ascii=False, indent=2)


def generate_dataset(ref_src: str, out_dir: str, n: int = 10, seed: Optional[int] = None, grid: int = 3, workers: Optional[int] = None,
                     device: int = 0, batch_panels: int = 4096, quiet: bool = True) -> Dict[str, Any]:
    """`python src/cli.py --out_dir … --n … --seed … --workers …` with the pixels made on the GPU.  Returns counters."""
    import time
    t0 = time.perf_counter()
    # the pool is created (forked) by the first next() below, before this process creates its CUDA context in DatasetWriter
    stream = emit_dataset(ref_src, out_dir, n, seed, grid, workers, quiet)
    first = next(stream, None)
    writer = None
    metas: List[Any] = [None] * n
    ok = 0
    try:
        em = first
        while em is not None:
            metas[em.index] = em.meta
            if em.ok:
                if writer is None:
                    writer = DatasetWriter(em.scene.W, em.scene.H, device=device, batch_panels=batch_panels)
                writer.add(em)
                ok += 1
            em = next(stream, None)
    finally:
        if writer is not None:
            writer.close()
    write_index(out_dir, metas)
    dt = time.perf_counter() - t0
    return dict(samples=n, ok=ok, failed=n - ok, seconds=dt, samples_per_s=n / dt if dt > 0 else 0.0,
                panels=writer.panels if writer else 0, files=writer.files if writer else 0,
                png_bytes=writer.png_bytes if writer else 0)


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description='Batched GPU drop-in for the reference src/cli.py')
    ap.add_argument('--ref_src', required=True, help="the reference's src/ directory (rules, sampling and layout logic run from there)")
    ap.add_argument('--out_dir', type=str, default='../data_251027_2w')
    ap.add_argument('--n', type=int, default=10)
    ap.add_argument('--grid', type=int, default=3)
    ap.add_argument('--seed', type=int, default=None)
    ap.add_argument('--workers', type=int, default=None)
    ap.add_argument('--device', type=int, default=0)
    a = ap.parse_args(argv)
    print(json.dumps(generate_dataset(a.ref_src, a.out_dir, a.n, a.seed, a.grid, a.workers, a.device)))


if __name__ == '__main__':
    main()
