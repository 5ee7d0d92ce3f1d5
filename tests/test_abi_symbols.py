"""The C-ABI library builds for sm_100a, loads without a GPU, and exports exactly what include/rig.h declares."""
import ctypes
import math
import os
import re

import numpy as np

from conftest import ROOT
from reasoning_image_generation_b200 import _abi


def declared_functions():
    txt = open(os.path.join(ROOT, 'include', 'rig.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(rig_[a-z_]+)\s*\(', txt)))


def test_every_declared_symbol_is_exported(built):
    lib = _abi.load()
    names = declared_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_abi.EXPORTS) == names


def test_every_declared_figure_symbol_is_exported(built):
    """include/rig_figure.h: the figure (matplotlib/Agg) path of the same library."""
    from reasoning_image_generation_b200 import figure
    lib = _abi.load()
    txt = open(os.path.join(ROOT, 'include', 'rig_figure.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    names = sorted(set(re.findall(r'\b(rig_figure_[a-z_]+)\s*\(', txt)))
    assert len(names) == 17
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(figure.FIGURE_EXPORTS) == names


def test_every_declared_png_symbol_is_exported(built):
    """include/rig_png.h: the device-side PNG writer of the same library."""
    from reasoning_image_generation_b200 import png
    lib = _abi.load()
    txt = open(os.path.join(ROOT, 'include', 'rig_png.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    names = sorted(set(re.findall(r'\b(rig_png_[a-z_]+)\s*\(', txt)))
    assert len(names) == 13
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(png.PNG_EXPORTS) == names


def test_every_declared_grid_symbol_is_exported(built):
    """include/rig_grid.h: the overview compositor of the same library."""
    from reasoning_image_generation_b200 import compose
    lib = _abi.load()
    txt = open(os.path.join(ROOT, 'include', 'rig_grid.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    names = sorted(set(re.findall(r'\b(rig_grid_[a-z_]+)\s*\(', txt)))
    assert len(names) == 7
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(compose.GRID_EXPORTS) == names


def test_strict_png_reader_agrees_with_libpng():
    """The tests' pure-zlib PNG reader decodes what cv2.imwrite (libpng) writes."""
    import cv2
    from reasoning_image_generation_b200.png import decode_png_strict
    rng = np.random.RandomState(0)
    img = rng.randint(0, 256, (37, 48, 3)).astype(np.uint8)
    img[5:20, 3:40] = 255
    ok, enc = cv2.imencode('.png', img[:, :, ::-1], [cv2.IMWRITE_PNG_STRATEGY, cv2.IMWRITE_PNG_STRATEGY_RLE])
    assert ok
    try:
        got = decode_png_strict(enc.tobytes())
    except AssertionError as e:
        if 'filter type' in str(e):
            return                       # libpng chose a filter this reader does not implement: nothing to compare
        raise
    assert (got == img).all()


def test_figure_create_without_device_fails_loudly(built):
    import torch
    if torch.cuda.is_available():
        return
    from reasoning_image_generation_b200 import figure
    with __import__('pytest').raises(figure.FigureError) as e:
        figure.FigureRenderer(512, 512)
    assert e.value.code in (_abi.RIG_E_NODEVICE, _abi.RIG_E_CUDA)


def test_create_without_device_fails_loudly_not_silently(built):
    """No CPU fallback: on a box without a usable GPU rig_create returns an error and a message."""
    import torch
    if torch.cuda.is_available():
        return
    lib = _abi.load()
    h = ctypes.c_void_p()
    rc = lib.rig_create(0, 512, 512, 16, 64, ctypes.byref(h))
    assert rc in (_abi.RIG_E_NODEVICE, _abi.RIG_E_CUDA) and not h.value
    assert b'CUDA' in lib.rig_last_error(None) or b'device' in lib.rig_last_error(None)


def test_bad_canvas_is_rejected_before_touching_cuda(built):
    lib = _abi.load()
    h = ctypes.c_void_p()
    assert lib.rig_create(0, 510, 512, 16, 64, ctypes.byref(h)) == _abi.RIG_E_INVALID
    assert lib.rig_create(0, 512, 0, 16, 64, ctypes.byref(h)) == _abi.RIG_E_INVALID


def test_shape_tables_equal_the_python_expressions_of_the_reference(built):
    """The unit tables of the device prologue are the libm values CPython computes
    (/root/reference/src/shapes.py:431-435,456-460,518-527,613-637)."""
    lib = _abi.load()
    size = lib.rig_debug_shape_tables(None, 0)
    buf = (ctypes.c_uint8 * size)()
    assert lib.rig_debug_shape_tables(buf, size) == size
    raw = bytes(buf)
    reg = np.frombuffer(raw, np.float64, 36, 0).reshape(3, 6, 2)
    heart = np.frombuffer(raw, np.float64, 120, 36 * 8).reshape(60, 2)
    arcs = np.frombuffer(raw, np.float64, 96, (36 + 120) * 8).reshape(48, 2)
    sin_tab = np.frombuffer(raw, np.float32, 451, (36 + 120 + 96) * 8)
    for i in range(5):
        a = math.radians(i * 72 - 90)
        assert (reg[0, i, 0], reg[0, i, 1]) == (math.cos(a), math.sin(a))
        a = math.radians(i * (360.0 / 5) - 90)
        assert (reg[1, i, 0], reg[1, i, 1]) == (math.cos(a), math.sin(a))
    for i in range(6):
        a = math.radians(i * (360.0 / 6) - 90)
        assert (reg[2, i, 0], reg[2, i, 1]) == (math.cos(a), math.sin(a))
    for i, t in enumerate(np.linspace(0, 2 * math.pi, num=60)):
        x = 16 * (math.sin(t) ** 3)
        y = 13 * math.cos(t) - 5 * math.cos(2 * t) - 2 * math.cos(3 * t) - math.cos(4 * t)
        assert (heart[i, 0], heart[i, 1]) == (x, y), i
    spans = [(math.pi, 3 * math.pi / 2), (3 * math.pi / 2, 2 * math.pi), (0, math.pi / 2), (math.pi / 2, math.pi)]
    for c, (a0, a1) in enumerate(spans):
        for j, th in enumerate(np.linspace(a0, a1, num=12)):
            assert (arcs[c * 12 + j, 0], arcs[c * 12 + j, 1]) == (math.cos(th), math.sin(th)), (c, j)
    want = np.array([np.float32('%.7f' % math.sin(math.radians(d))) for d in range(451)])
    assert (sin_tab.view(np.uint32) == want.view(np.uint32)).all()
