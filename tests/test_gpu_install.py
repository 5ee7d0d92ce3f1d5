"""install(): the one-panel-per-call drop-in seam, under the reference's own process / thread structure:
`cli.py` forks pool workers (cli.py:85,100) and every worker renders from a NON-MAIN DAEMON THREAD
(generator.py:192-197, `_run_with_timeout`).  librig must therefore create its CUDA context lazily, in the process and
thread that first renders, never at import.  The forking parent must not hold a CUDA context itself (fork + CUDA is
undefined), so the structure is reproduced in a fresh interpreter."""
import hashlib
import json
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

from conftest import ROOT, restore_rng

pytestmark = pytest.mark.gpu

CHILD = r'''
import hashlib, json, os, random, sys, threading
from concurrent.futures import ProcessPoolExecutor
sys.path.insert(0, %(root)r)
import reasoning_image_generation_b200 as rig          # imported BEFORE the fork, as cli.py imports generator.py


class Stub:                                            # stands for generator.RPMGenerator
    def _render_elements_to_canvas(self, W, H, bg_color, elements, use_grid=False, grid_size=3):
        raise NotImplementedError


rig.install(Stub)
CASES = json.load(open(%(cases)r))


def make_sample(k):
    """cli._make_sample -> RPMGenerator.generate_sample -> _run_with_timeout: a daemon thread inside the pool worker."""
    res, exc = [None], [None]

    def target():
        try:
            c = CASES[k]
            st = c.get('rng_state')
            if st:
                random.setstate((st[0], tuple(st[1]), st[2]))
            img = Stub()._render_elements_to_canvas(c['W'], c['H'], tuple(c['bg']), c['elements'], c['use_grid'], c['grid_size'])
            res[0] = (hashlib.sha256(img.tobytes()).hexdigest(), os.getpid(), threading.current_thread() is threading.main_thread())
        except Exception as e:
            exc[0] = repr(e)
    t = threading.Thread(target=target)
    t.daemon = True
    t.start()
    t.join(120)
    return res[0], exc[0]


if __name__ == '__main__':
    with ProcessPoolExecutor(max_workers=2) as exe:      # fork start method, like the reference on Linux
        out = list(exe.map(make_sample, range(len(CASES))))
    print(json.dumps(out))
'''


def test_install_in_forked_workers_on_daemon_threads(golden_seq, built, tmp_path):
    panels, pix = golden_seq
    pick = [0, 3, 5, 8, 13, 21]
    cases = tmp_path / 'cases.json'
    cases.write_text(json.dumps([panels[k] for k in pick]))
    script = tmp_path / 'child.py'
    script.write_text(CHILD % dict(root=ROOT, cases=str(cases)))
    env = dict(os.environ)
    env.pop('PYTHONSTARTUP', None)
    p = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    out = json.loads(p.stdout.strip().splitlines()[-1])
    pids = set()
    for k, (res, exc) in zip(pick, out):
        assert exc is None, exc
        sha, pid, on_main = res
        assert sha == hashlib.sha256(pix[k].tobytes()).hexdigest(), k
        assert not on_main
        pids.add(pid)
    assert os.getpid() not in pids


def test_install_from_a_daemon_thread_in_this_process(golden_seq, built):
    from reasoning_image_generation_b200 import install

    class Stub:
        pass

    install(Stub)
    panels, pix = golden_seq
    got = {}

    def target():
        for k in (1, 4):
            c = panels[k]
            restore_rng(c)
            els = [dict(e) for e in c['elements']]
            got[k] = Stub()._render_elements_to_canvas(c['W'], c['H'], tuple(c['bg']), els, c['use_grid'], c['grid_size'])
            assert els == [dict(e) for e in c['elements']]           # elements untouched
    t = threading.Thread(target=target, daemon=True)
    t.start()
    t.join(120)
    assert not t.is_alive() and sorted(got) == [1, 4]
    for k, img in got.items():
        assert img.dtype == np.uint8 and img.shape == (512, 512, 3) and (img == pix[k]).all()
