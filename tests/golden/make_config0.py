"""Freezes BASELINE.json configs[0] -- `src/cli.py --n 1000 --seed 0` -- as rendered by the UNMODIFIED live reference.

Run in the build container only (needs /root/reference/src and cv2):

    python tests/golden/make_config0.py [n]

Every sample i is generated exactly as cli._make_sample does (GenConfig(seed=0 + i), RPMGenerator, sample id i), calling
`_generate_sample_impl` under a 10 s alarm instead of the 30 s watchdog thread (the samples that hang are the reference's
`while(1)` bug, SURVEY B2: they never finish, so the set of lost samples does not depend on the limit).  Written:

  config0.json.gz   per sample: ok, rule, correct_index, sha256 of the normalised meta.json, the overview's sha256 and, per
                    render call (generator.py:66), the element fields the renderer reads, use_grid, the colours drawn for
                    colourless elements (B8) and the sha256 of the returned BGR pixels; per saved file its panel; the panels the
                    overview shows (compose_grid's states / candidates, layout.py:98-174)
Nothing at test time reads /root/reference: tests/test_config0.py replays the element lists on the GPU (-m gpu) and, where
the reference is present, re-runs the emit stage of reasoning_image_generation_b200.driver against the same records (CPU).
"""
import gzip
import hashlib
import json
import multiprocessing as mp
import os
import random
import signal
import sys
import tempfile

import numpy as np

REF = '/root/reference/src'
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
RENDER_FIELDS = ('kind', 'size', 'fill', 'stroke_width', 'center', 'color', 'angle')


def normalise_meta(meta, out_dir):
    """meta.json without what changes from run to run: timestamps, tracebacks, the output directory prefix."""
    def walk(v):
        if isinstance(v, dict):
            return {k: walk(x) for k, x in v.items() if k not in ('timestamp', 'generation_time', 'traceback')}
        if isinstance(v, (list, tuple)):
            return [walk(x) for x in v]
        if isinstance(v, str):
            return v.replace(out_dir, '<out>')
        if isinstance(v, np.generic):
            return v.item()
        return v
    return walk(meta)


def meta_digest(meta, out_dir):
    return hashlib.sha256(json.dumps(normalise_meta(meta, out_dir), sort_keys=True, ensure_ascii=False, default=str).encode()).hexdigest()


def render_view(el):
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


class Stuck(BaseException):
    """BaseException: the reference wraps its rule handlers in `except Exception` (generator.py:372-385,476-484), which must
    not swallow the alarm."""


def _alarm(*_):
    raise Stuck()


_state = {}


def _init():
    sys.path.insert(0, REF)
    import logging
    logging.disable(logging.CRITICAL)
    import generator as refgen
    from config import GenConfig
    _state['refgen'], _state['GenConfig'] = refgen, GenConfig
    _state['tmp'] = tempfile.mkdtemp(prefix='rig_config0_')
    cap = _state['cap'] = {'panels': [], 'by_id': {}, 'files': [], 'keep': []}
    orig = refgen.RPMGenerator._render_elements_to_canvas
    orig_save = refgen.save_image

    def spy(self, W, H, bg_color, elements, use_grid=False, grid_size=3):
        before = random.getstate()
        out = orig(self, W, H, bg_color, elements, use_grid, grid_size)
        after = random.getstate()
        random.setstate(before)
        drawn = [[int(random.uniform(30, 220)) for _ in range(3)] if e.get('color') is None else None for e in elements]
        assert random.getstate() == after, 'colour draws do not account for the RNG use of the render call'
        cap['by_id'][id(out)] = len(cap['panels'])
        cap['keep'].append(out)                          # keeps id() unique for the sample
        p = dict(elements=[render_view(e) for e in elements], use_grid=bool(use_grid), sha=hashlib.sha256(out.tobytes()).hexdigest())
        if any(d is not None for d in drawn):
            p['drawn'] = drawn
        cap['panels'].append(p)
        return out

    def save_spy(im, path):
        k = cap['by_id'].get(id(im))
        if k is not None:
            cap['files'].append([os.path.relpath(path, _state['tmp']), k])
        else:
            cap['grid_sha'] = hashlib.sha256(np.ascontiguousarray(im).tobytes()).hexdigest()
            cap['grid_file'] = os.path.relpath(path, _state['tmp'])
            cap['grid_shape'] = list(im.shape)
        return orig_save(im, path)

    orig_compose = refgen.compose_grid

    def compose_spy(*a, **kw):
        states, candidates = kw['states'], kw['candidates']
        cap['grid_states'] = [cap['by_id'][id(s['state_img'])] for s in states]
        cap['grid_options'] = [cap['by_id'][id(c['img'])] for c in candidates[:kw['num_options']]]
        return orig_compose(*a, **kw)

    refgen.RPMGenerator._render_elements_to_canvas = spy
    refgen.save_image = save_spy
    refgen.compose_grid = compose_spy
    signal.signal(signal.SIGALRM, _alarm)


def one(i):
    refgen, GenConfig, cap, tmp = _state['refgen'], _state['GenConfig'], _state['cap'], _state['tmp']
    cap['panels'], cap['by_id'], cap['files'], cap['keep'] = [], {}, [], []
    cap.pop('grid_sha', None)
    cfg = GenConfig(out_dir=tmp, grid_size=3, seed=0 + i)
    gen = refgen.RPMGenerator(cfg)
    signal.alarm(10)
    try:
        meta = gen._generate_sample_impl(i)
    except Stuck:
        return dict(index=i, ok=False, why='B2 while(1)')
    except Exception as e:                               # generate_sample logs and returns None (generator.py:226-230)
        return dict(index=i, ok=False, why=f'{type(e).__name__}: {e}')
    finally:
        signal.alarm(0)
    import cv2
    q = cv2.imread(os.path.join(tmp, 'samples', f'sample_{i:06d}', 'query.png'), cv2.IMREAD_UNCHANGED)
    return dict(index=i, ok=True, rule=meta['rule'], correct_index=meta['correct_index'], meta_sha=meta_digest(meta, tmp),
                n_states=len(meta['sequence']), panels=cap['panels'], files=cap['files'], grid_sha=cap.get('grid_sha'),
                grid_file=cap.get('grid_file'), grid_shape=cap.get('grid_shape'), grid_states=cap['grid_states'], grid_options=cap['grid_options'],
                query_sha=hashlib.sha256(np.ascontiguousarray(q).tobytes()).hexdigest() if q is not None else None)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    with mp.get_context('fork').Pool(os.cpu_count(), initializer=_init) as pool:
        recs = pool.map(one, range(n), chunksize=4)
    ok = [r for r in recs if r['ok']]
    import cv2
    out = dict(seed=0, n=n, cv2=cv2.__version__, numpy=np.__version__, samples=recs)
    path = os.path.join(HERE, 'config0.json.gz')
    with gzip.GzipFile(path, 'wb', mtime=0) as f:
        f.write(json.dumps(out, ensure_ascii=False, separators=(',', ':')).encode())
    print(f'{len(ok)}/{n} samples ok, {sum(len(r["panels"]) for r in ok)} panels, '
          f'{sum(1 for r in ok for p in r["panels"] if "drawn" in p)} with colourless elements -> {path} ({os.path.getsize(path)} bytes)')


if __name__ == '__main__':
    main()
