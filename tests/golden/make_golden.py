"""Generates the parity fixtures in this directory FROM THE LIVE REFERENCE.  Run in the build container only
(needs /root/reference/src and cv2); the fixtures it writes are committed so that nothing at test time
reads /root/reference.

    python tests/golden/make_golden.py

Writes
  seq_panels.json / seq_panels.npz   panels captured at RPMGenerator._render_elements_to_canvas
                                     (/root/reference/src/generator.py:66) while running
                                     _generate_sample_impl for seeds 777..
  kinds_panels.json / kinds_panels.npz  hand-made element lists (every kind x angle x stroke x fill,
                                     wrap-around placements, colourless elements) rendered by calling the
                                     reference's _render_elements_to_canvas directly
  dedup_keys.json                    reference _elements_set equality decisions (generator.py:148-177)
Pixels are stored as uint8[H,W,3] BGR exactly as the reference returns them.
"""
import json
import os
import random
import signal
import sys
import tempfile

import numpy as np

REF = '/root/reference/src'
HERE = os.path.dirname(os.path.abspath(__file__))


def jsonable(v):
    if isinstance(v, dict):
        return {k: jsonable(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [jsonable(x) for x in v]
    if isinstance(v, (np.integer,)):
        return int(v)
    if isinstance(v, (np.floating,)):
        return float(v)
    return v


def main():
    sys.path.insert(0, REF)
    import logging
    logging.disable(logging.CRITICAL)
    from config import GenConfig
    import generator as refgen

    captured = []
    orig = refgen.RPMGenerator._render_elements_to_canvas

    def spy(self, W, H, bg_color, elements, use_grid=False, grid_size=3):
        state = random.getstate()
        out = orig(self, W, H, bg_color, elements, use_grid, grid_size)
        captured.append(dict(W=W, H=H, bg=list(bg_color), use_grid=bool(use_grid), grid_size=grid_size,
                             elements=jsonable(elements), rng_draws=sum(1 for e in elements if e.get('color') is None),
                             _rng_state=state, _pix=out.copy()))
        return out

    refgen.RPMGenerator._render_elements_to_canvas = spy

    class Stuck(Exception):
        pass

    def on_alarm(*_):
        raise Stuck()

    signal.signal(signal.SIGALRM, on_alarm)
    tmp = tempfile.mkdtemp(prefix='rig_golden_')
    seq = []
    n_samples = 0
    for i in range(40):
        cfg = GenConfig(out_dir=tmp, seed=777 + i)
        gen = refgen.RPMGenerator(cfg)
        captured.clear()
        signal.alarm(10)
        try:
            meta = gen._generate_sample_impl(i)
        except Stuck:
            print('sample', i, 'hit the reference while(1) bug (SURVEY B2); skipped')
            continue
        finally:
            signal.alarm(0)
        n_samples += 1
        rule = meta.get('rule_leaf') or meta.get('category_path') if isinstance(meta, dict) else None
        for c in captured:
            c['sample'] = i
            c['rule'] = rule
            seq.append(dict(c))
    # keep the fixture small: every 2nd panel, but all panels that have >=4 elements or a colourless element
    keep = [c for k, c in enumerate(seq) if k % 3 == 0 or len(c['elements']) >= 4 or c['rng_draws']]
    print('captured', len(seq), 'panels from', n_samples, 'samples; keeping', len(keep))
    dump('seq_panels', keep)

    # ---- hand-made coverage set, rendered by the reference itself
    gen = refgen.RPMGenerator(GenConfig(out_dir=tmp, seed=1))
    kinds = ['square', 'circle', 'triangle', 'diamond', 'star', 'pentagon', 'hexagon', 'plus', 'heart',
             'crescent', 'rounded_square']
    rng = random.Random(20251018)
    cases = []
    for kind in kinds:
        for stroke in (1, 2, 3):
            for fill in (True, False):
                angle = rng.choice([0, 30, 45, 60, 90, 135, 180, 225, 270, 315, 7, 353])
                size = rng.choice([86, 102, 117, 128, 141, 170])
                cases.append(([dict(kind=kind, size=size, fill=fill, stroke_width=stroke,
                                    center=(rng.randint(90, 420), rng.randint(90, 420)), angle=float(angle),
                                    color=(rng.randint(30, 220), rng.randint(30, 220), rng.randint(30, 220)))], False))
    # wrap-around: centres near / beyond the borders and corners
    for kind in kinds:
        for (cx, cy) in [(5, 250), (508, 250), (250, 3), (250, 510), (4, 6), (509, 507), (-60, 200), (600, 300),
                         (256, -80), (300, 700), (-170, -170)]:
            cases.append(([dict(kind=kind, size=rng.choice([102, 128, 150]), fill=rng.random() < 0.7,
                                stroke_width=rng.randint(1, 3), center=(cx, cy),
                                angle=float(rng.choice([0, 45, 90, 135, 180, 30, 60])),
                                color=(rng.randint(30, 220), rng.randint(30, 220), rng.randint(30, 220)))], False))
    # grid panels, overlapping multi-element panels, float centres, colourless elements (B8)
    for k in range(40):
        n = rng.randint(1, 8)
        els = []
        for _ in range(n):
            els.append(dict(kind=rng.choice(kinds), size=rng.choice([102, 110, 128, 160]), fill=rng.random() < 0.67,
                            stroke_width=rng.randint(1, 3),
                            center=(rng.uniform(0, 512), rng.uniform(0, 512)) if k % 3 == 0 else (rng.randint(0, 512), rng.randint(0, 512)),
                            angle=float(rng.choice([0, 45, 90, 135, 180, 30, 60, 300])),
                            color=None if (k % 5 == 0 and rng.random() < 0.5) else (rng.randint(30, 220), rng.randint(30, 220), rng.randint(30, 220))))
        cases.append((els, k % 2 == 0))
    cases.append(([], False))
    cases.append(([], True))
    captured.clear()
    random.seed(424242)
    for els, ug in cases:
        gen._render_elements_to_canvas(512, 512, (255, 255, 255), els, use_grid=ug)
    for c in captured:
        c['sample'] = -1
        c['rule'] = None
    print('hand-made panels', len(captured))
    dump('kinds_panels', [dict(c) for c in captured])

    # ---- in-sample option de-dup decisions (a10)
    S = refgen.RPMGenerator._elements_set
    pairs = []
    pool = [c['elements'] for c in keep[:60]]
    for a in range(len(pool)):
        for b in range(a, min(a + 6, len(pool))):
            pairs.append(dict(a=a, b=b, equal=bool(S(pool[a]) == S(pool[b]))))
    extra = []
    base = [dict(kind='square', size=100, fill=True, stroke_width=2, center=(10, 20), angle=0.0,
                 bbox=(1, 2, 3, 4), flip={'h': False, 'v': True}, color=(1, 2, 3))]
    variants = [
        base,
        [dict(base[0], timestamp='x', temp_id=7)],
        [dict(base[0], center=[10, 20])],
        [dict(base[0], center=(10.0, 20.0))],
        [dict(base[0], size=np.int64(100))],
        [dict(base[0], flip={'v': True, 'h': False})],
        [dict(base[0]), dict(base[0])],
        [dict(base[0], color=None)],
        [dict(base[0], angle=1.0)],
        [],
    ]
    for a in range(len(variants)):
        for b in range(len(variants)):
            extra.append(dict(a=a, b=b, equal=bool(S(variants[a]) == S(variants[b]))))
    with open(os.path.join(HERE, 'dedup_keys.json'), 'w') as f:
        json.dump(dict(pool=jsonable(pool), pairs=pairs, variants=jsonable(variants), variant_pairs=extra,
                       note='variants[4].size was np.int64(100) when the reference evaluated it'), f)


def dump(name, panels):
    pix = np.stack([c.pop('_pix') for c in panels])
    states = [c.pop('_rng_state') for c in panels]
    for c, st in zip(panels, states):
        c['rng_state'] = [st[0], list(st[1]), st[2]] if c['rng_draws'] else None
    with open(os.path.join(HERE, name + '.json'), 'w') as f:
        json.dump(panels, f)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), pixels=pix)
    print(name, pix.shape, os.path.getsize(os.path.join(HERE, name + '.npz')) // 1024, 'KiB')


if __name__ == '__main__':
    main()
