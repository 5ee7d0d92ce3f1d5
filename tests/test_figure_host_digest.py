"""The host path pipeline's output (layers + 24.8 fixed-point edges, what is uploaded to the GPU) pinned by digest.

The pixel-level truth is the Agg oracle (test_figure_oracle.py, test_figure_emulation.py); this test is the cheap guard for
refactors of csrc/rig_figure_path.cpp: its arithmetic must stay operation for operation what matplotlib's converters and
Agg's stroker do, so ANY change of an edge coordinate or of the edge order is a finding.  The digests were written by
the build that was checked edge by edge against the previous implementation (tests/golden/figure_host_digests.json; the scene
generators are seeded).  They depend on the host libm (sqrt, atan2, cos, sin, acos): the image this repo runs in."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))


def _digest(layers, edges):
    return hashlib.sha256(np.ascontiguousarray(layers).tobytes() + np.ascontiguousarray(edges).tobytes()).hexdigest()


def test_host_pipeline_output_is_unchanged(built):
    from figure_stress_scene import stress_scene
    from reasoning_image_generation_b200.figure import host_edges
    from reasoning_image_generation_b200.synth_figure import synthetic_figure_scene
    want = json.load(open(os.path.join(HERE, 'golden', 'figure_host_digests.json')))
    scenes = {'stress': stress_scene(), 'single': synthetic_figure_scene(1500, 512, 512, seed=77, kind='single'),
              'multi': synthetic_figure_scene(1500, 512, 512, seed=77, kind='multi')}
    for name, sc in scenes.items():
        layers, edges = host_edges(sc)
        assert (len(layers), len(edges)) == (want[name]['layers'], want[name]['edges']), name
        assert _digest(layers, edges) == want[name]['sha256'], name


def test_stress_scene_renders_like_the_oracle(built):
    """The same stress scene through the CPU emulation of k_figure against the Agg oracle: every branch the digest pins is also
    pixel-checked."""
    import ctypes
    from conftest import ROOT
    from figure_stress_scene import stress_scene
    from oracle.agg_oracle import AggOracle
    from test_figure_emulation import emul_render
    lib = ctypes.CDLL(os.path.join(ROOT, 'tests', 'cpu_emul', 'libfigure_emul.so'))
    lib.figure_emul_render.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    lib.figure_emul_render.restype = None
    sc = stress_scene()
    ref = AggOracle().render(sc, sc.W, sc.H)
    got = emul_render(lib, sc, 2)
    bad = [i for i in range(sc.n_images) if not (got[i] == ref[i]).all()]
    assert not bad, 'images that differ from the oracle: %s' % bad[:10]
