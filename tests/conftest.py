import ctypes
import json
import os
import random
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def built():
    """Everything native, built once: librig.so (nvcc), the C oracle (gcc), the CPU emulation (g++)."""
    import __graft_entry__ as ge
    ge.build()
    return True


@pytest.fixture(scope='session')
def oracle_lib(built):
    from oracle import ref_render
    lib = ctypes.CDLL(ref_render.build_c_oracle())
    return lib


@pytest.fixture(scope='session')
def emul_lib(built):
    from reasoning_image_generation_b200._abi import RigScene
    lib = ctypes.CDLL(os.path.join(ROOT, 'tests', 'cpu_emul', 'libemul.so'))
    lib.emul_render.argtypes = [ctypes.POINTER(RigScene), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint32), ctypes.c_char_p, ctypes.c_size_t]
    lib.emul_render.restype = ctypes.c_int
    lib.emul_shape_tables.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    lib.emul_shape_tables.restype = ctypes.c_size_t
    return lib


def load_golden(name):
    panels = json.load(open(os.path.join(GOLDEN, name + '.json')))
    pix = np.load(os.path.join(GOLDEN, name + '.npz'))['pixels']
    assert len(panels) == len(pix)
    return panels, pix


def restore_rng(panel):
    """Panels with colourless elements were rendered by the reference from this `random` state."""
    st = panel.get('rng_state')
    if st:
        random.setstate((st[0], tuple(st[1]), st[2]))


@pytest.fixture(scope='session')
def golden_seq():
    return load_golden('seq_panels')


@pytest.fixture(scope='session')
def golden_kinds():
    return load_golden('kinds_panels')


def emul_render_scene(lib, scene, band_rows=8, layout=1):
    out = np.zeros((scene.n_panels, scene.H, scene.W, 3), np.uint8)
    err = ctypes.create_string_buffer(256)
    nrec = ctypes.c_uint32()
    st = scene.as_struct()
    rc = lib.emul_render(ctypes.byref(st), scene.W, scene.H, band_rows, layout, out.ctypes.data, ctypes.byref(nrec), err, 256)
    return rc, out, nrec.value, err.value.decode()
