"""BASELINE.json configs[0] (`src/cli.py --n 1000 --seed 0`) as a parity run.

tests/golden/config0.json.gz was recorded from the UNMODIFIED live reference (tests/golden/make_config0.py): for each of the
1 000 samples its element lists per render call, the sha256 of every rendered panel and of the overview, the files saved and
the sha256 of the normalised meta.json; 33 samples are lost to the reference's while(1) bug.

CPU:  the fixture is consistent; the oracle reproduces a spread of its panels; where /root/reference is present the emit
      stage of `driver` (the reference's rule logic with recording seams) reproduces every sample's element lists, colours
      drawn, file plan and labels/meta bit for bit and loses exactly the same 33 samples -- without waiting for a watchdog.
GPU:  all ~7.4 k panels rendered in ONE batch equal the recorded digests; the owner stage (render + compose + device PNG +
      file layout) writes files that decode to the recorded pixels.
"""
import gzip
import hashlib
import json
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN

REF_SRC = '/root/reference/src'


@pytest.fixture(scope='module')
def config0():
    return json.loads(gzip.open(os.path.join(GOLDEN, 'config0.json.gz')).read())


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def resolved(panel):
    """Element dicts with the colours the reference drew for colourless elements filled in."""
    drawn = panel.get('drawn')
    out = []
    for k, e in enumerate(panel['elements']):
        e = dict(e)
        if e.get('color') is None:
            e['color'] = drawn[k]
        out.append(e)
    return out


def test_fixture_shape(config0):
    s = config0['samples']
    assert config0['n'] == len(s) == 1000 and [r['index'] for r in s] == list(range(1000))
    ok = [r for r in s if r['ok']]
    assert len(ok) == 967 and all(r['why'] == 'B2 while(1)' for r in s if not r['ok'])
    assert sum(len(r['panels']) for r in ok) == 7431
    for r in ok:
        L = r['n_states']
        assert L in (4, 6) and len(r['panels']) == L + 3 and len(r['grid_states']) == L - 1 and len(r['grid_options']) == 4
        names = [os.path.basename(f) for f, _ in r['files']]
        assert names == [f'state_{i}.png' for i in range(L)] + ['proto_true_next.png', 'option_1.png', 'option_2.png', 'option_3.png']
        assert r['grid_shape'] == [296 if L == 4 else 216, 512, 3]     # cell = 472 // (L - 1 + 1) capped by the height


def test_oracle_reproduces_config0_panels(config0, built):
    from oracle import ref_render
    be = ref_render.CBackend()
    picks = [(r, p) for r in config0['samples'] if r['ok'] for p in r['panels']]
    n = 0
    for r, p in picks[::37] + [(r, p) for r, p in picks if 'drawn' in p][:10]:
        got = ref_render.render_panel(512, 512, (255, 255, 255), resolved(p), p['use_grid'], 3, backend=be)
        assert sha(got) == p['sha'], r['index']
        n += 1
    assert n > 200


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason='the reference lives only in the build container')
def test_emit_stage_equals_reference(config0, tmp_path):
    sys.path.insert(0, GOLDEN)
    from make_config0 import meta_digest
    from reasoning_image_generation_b200 import driver
    out = str(tmp_path)
    seen = 0
    for em in driver.emit_dataset(REF_SRC, out, 1000, 0, 3, workers=min(8, os.cpu_count() or 1)):
        r = config0['samples'][em.index]
        assert em.ok == r['ok'], (em.index, r.get('why'))
        seen += 1
        if not em.ok:
            assert em.meta is None                       # generate_sample's result for a lost sample (generator.py:221-230)
            continue
        assert meta_digest(em.meta, out) == r['meta_sha'], em.index
        assert (em.meta['rule'], em.meta['correct_index']) == (r['rule'], r['correct_index'])
        assert json.loads(json.dumps([dict(elements=e, use_grid=g) for e, g in em.panels])) == \
            [dict(elements=p['elements'], use_grid=p['use_grid']) for p in r['panels']], em.index
        assert [[os.path.relpath(p, out), k] for p, k in em.files] == r['files']
        assert os.path.relpath(em.grid_path, out) == r['grid_file'] and em.grid_states == r['grid_states'] and em.grid_options == r['grid_options']
        want = [e['color'] for p in r['panels'] for e in resolved(p)]
        assert em.scene.rgb.reshape(-1, 3).tolist() == [[min(255, max(0, int(c))) for c in w] for w in want], em.index
        assert os.path.isfile(os.path.join(out, 'samples', f'sample_{em.index:06d}', 'meta.json'))
    assert seen == 1000
    assert not os.path.exists(os.path.join(out, 'samples', 'sample_000004')) or config0['samples'][4]['ok']


def _emissions(config0, out_dir, limit=None):
    """Emission objects rebuilt from the fixture (the GPU box has no reference to run the emit stage with)."""
    from reasoning_image_generation_b200.driver import Emission, GridRef, Tape
    ems = []
    for r in config0['samples']:
        if not r['ok']:
            continue
        t = Tape(512, 512, (255, 255, 255), 3)
        for p in r['panels']:
            t.builder.add_panel(resolved(p), p['use_grid'])
        t.files = [(os.path.join(out_dir, f), k) for f, k in r['files']]
        t.grid = GridRef(r['grid_states'], r['grid_options'])
        t.grid_path = os.path.join(out_dir, r['grid_file'])
        t.query_path = os.path.join(out_dir, 'samples', f'sample_{r["index"]:06d}', 'query.png')
        t.layout_key = (512, 512, len(r['grid_states']), 4, 20, 20, True, True, (255, 255, 255))
        for d in {os.path.dirname(t.grid_path), os.path.dirname(t.query_path)}:
            os.makedirs(d, exist_ok=True)
        ems.append(Emission(r['index'], {'id': r['index']}, t))
        if limit and len(ems) >= limit:
            break
    return ems


@pytest.mark.gpu
def test_all_config0_panels_in_one_batch(config0, built):
    from reasoning_image_generation_b200 import Renderer, SceneBuilder
    sb = SceneBuilder(512, 512)
    want = []
    for r in config0['samples']:
        if r['ok']:
            for p in r['panels']:
                sb.add_panel(resolved(p), p['use_grid'])
                want.append(p['sha'])
    sc = sb.finish()
    assert sc.n_panels == 7431
    r = Renderer(512, 512, device=0, max_panels=sc.n_panels, max_elements=sc.n_elements + 8)
    out = r.render(sc, 'bgr')
    r.close()
    bad = [i for i in range(sc.n_panels) if sha(out[i]) != want[i]]
    assert not bad, (len(bad), bad[:10])


@pytest.mark.gpu
def test_owner_stage_writes_the_reference_files(config0, built, tmp_path):
    import cv2
    from reasoning_image_generation_b200.driver import DatasetWriter
    out = str(tmp_path)
    ems = _emissions(config0, out, limit=160)
    w = DatasetWriter(512, 512, device=0, batch_panels=512)      # several batches, mixed L = 4 / L = 6 layouts
    for em in ems:
        w.add(em)
    w.close()
    assert w.samples == len(ems) and w.panels == sum(em.scene.n_panels for em in ems)
    by_index = {r['index']: r for r in config0['samples']}
    for em in ems:
        r = by_index[em.index]
        for (path, k) in em.files:
            img = cv2.imread(path, cv2.IMREAD_UNCHANGED)         # files hold RGB; imread hands back BGR, like the reference's arrays
            assert img is not None and sha(img) == r['panels'][k]['sha'], path
        g = cv2.imread(em.grid_path, cv2.IMREAD_UNCHANGED)
        assert g is not None and list(g.shape) == r['grid_shape'] and sha(g) == r['grid_sha'], em.grid_path
        q = cv2.imread(em.query_path, cv2.IMREAD_UNCHANGED)
        assert q is not None and sha(q) == r['query_sha'], em.query_path
