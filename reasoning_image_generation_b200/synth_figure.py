"""synth_figure.py -- synthetic multigraph figure scenes of the shape BASELINE.json configs 2 and 3 name.

The reference's multigraph generator needs matplotlib + shapely (absent here), so bench.py and the parity tests use scenes
that follow its parameter distributions:
  * single-shape variants (multigraph_generation/generator.py:105-128,198-228; single_variants.py:233-396,398-633): one base
    patch of size 3.0 * 1.3 data units -- circle, ellipse (x 0.6, random rotation), rectangle (30 %: rounded corners),
    regular 3..8-gon, wedge -- outline only, black, alpha 0.9, 1.5-2.0 pt; 70 %: 1-6 internal lines (radial / grid / polygon /
    chords; colour cycle, alpha 0.8-0.9, 0.6-1.4 pt, projecting caps, round joins); 30 %: 1-3 occluders (disc or rectangle)
    drawn as white fills plus black boundary segments with butt caps;
  * multi-shape combinations (generator.py:230-248; multi_combinator.py:685,774,1097): 2-3 base patches of size
    U(2.2, 3.2) * 1.3, one line width for all, 40 % rotated; nested (scale 0.4^i about the common centre), adjacent
    (touching side by side) or intersecting (centres closer than the radii).
Figure geometry: 8 in x 8 in at dpi = W / 8, data window (-5, 5)^2 in the default subplot slot (figure_paths.data_to_display).
Image i of a scene is seeded with random.Random(seed * 1000003 + i), so shards of a scene can be generated independently.
"""
from __future__ import annotations

import math
import random
from typing import Any, Dict, List, Tuple

import numpy as np

from . import figure_paths as fp
from .figure import FigureScene, FigureSceneBuilder

TAB10 = ((0x1f, 0x77, 0xb4), (0xff, 0x7f, 0x0e), (0x2c, 0xa0, 0x2c), (0xd6, 0x27, 0x28), (0x94, 0x67, 0xbd),
         (0x8c, 0x56, 0x4b), (0xe3, 0x77, 0xc2), (0x7f, 0x7f, 0x7f), (0xbc, 0xbd, 0x22), (0x17, 0xbe, 0xcf))
GLOBAL_SCALE = 1.3


class _Shape:
    """A base patch: its path, its patch transform (data space) and a boundary radius function for the decorations."""

    def __init__(self, kind, path, patch_affine, radius_at):
        self.kind, self.path, self.patch_affine, self.radius_at = kind, path, patch_affine, radius_at


def _ray_polygon(pts: np.ndarray):
    """Distance from the origin to the polygon boundary along angle t (polygon is star-shaped about the origin)."""
    def f(t):
        dx, dy = math.cos(t), math.sin(t)
        best = 0.0
        n = len(pts)
        for i in range(n):
            ax_, ay_ = pts[i]
            bx_, by_ = pts[(i + 1) % n]
            ex, ey = bx_ - ax_, by_ - ay_
            den = dx * ey - dy * ex
            if abs(den) < 1e-12:
                continue
            s = (ax_ * ey - ay_ * ex) / den
            u = (ax_ * dy - ay_ * dx) / den
            if s > 0 and -1e-9 <= u <= 1 + 1e-9:
                best = max(best, s)
        return best
    return f


def _base_shape(rng: random.Random, size: float, center=(0.0, 0.0)) -> _Shape:
    kind = rng.choice(('circle', 'ellipse', 'rectangle', 'regular_polygon', 'sector'))
    cx, cy = center
    if kind == 'circle':
        r = size / 2
        aff = fp.compose(fp.affine_scale(r, r), fp.affine_translate(cx, cy))
        return _Shape(kind, fp.unit_circle(), aff, lambda t: r)
    if kind == 'ellipse':
        a, b, ang = size / 2, size * 0.6 / 2, rng.uniform(-180, 180)
        aff = fp.compose(fp.affine_scale(a, b), fp.affine_rotate_deg(ang), fp.affine_translate(cx, cy))
        th = math.radians(ang)

        def rad(t, a=a, b=b, th=th):
            u = t - th
            return a * b / math.hypot(b * math.cos(u), a * math.sin(u))
        return _Shape(kind, fp.unit_circle(), aff, rad)
    if kind == 'rectangle':
        w, h = size, size * 0.6
        if rng.random() < 0.3:
            rc = rng.uniform(0.0, 0.5)
            path = fp.round_box(cx - w / 2, cy - h / 2, w, h, pad=0.0, rounding_size=max(rc, 1e-3))
            aff = fp.affine_identity()
        else:
            path = fp.unit_rectangle()
            aff = fp.compose(fp.affine_scale(w, h), fp.affine_translate(cx - w / 2, cy - h / 2))
        pts = np.array([[-w / 2, -h / 2], [w / 2, -h / 2], [w / 2, h / 2], [-w / 2, h / 2]])
        return _Shape(kind, path, aff, _ray_polygon(pts))
    if kind == 'regular_polygon':
        n, r = rng.randint(3, 8), size / 2
        v, c = fp.unit_regular_polygon(n)
        aff = fp.compose(fp.affine_scale(r, r), fp.affine_translate(cx, cy))
        return _Shape(kind, (v, c), aff, _ray_polygon(v[:-1] * r))
    r, t1, t2 = size / 2, rng.uniform(0, 180), rng.uniform(90, 360)
    v, c = fp.wedge_patch((cx, cy), r, t1, t2)
    return _Shape(kind, (v, c), fp.affine_identity(), lambda t: r * 0.7)


def _line(points, to_disp, lw_px, color, alpha, clip, cap='projecting', join='round', closed=False) -> Dict[str, Any]:
    v, c = fp.polyline(points, closed)
    return dict(vertices=v, codes=c, affine=to_disp, linewidth=lw_px, cap=cap, join=join, edgecolor=tuple(x / 255.0 for x in color) + (alpha,),
                facecolor=None, clip=clip)


def _decorations(rng: random.Random, sh: _Shape, center, to_disp, px_per_pt, clip) -> List[Dict[str, Any]]:
    out = []
    style = rng.choice(('radial', 'grid', 'random', 'polygon'))
    n = rng.randint(1, 6)
    cx, cy = center
    cyc = rng.randrange(10)

    def colour():
        nonlocal cyc
        col = TAB10[cyc % 10]
        cyc += 1
        return col
    if style == 'radial':
        for k in range(n):
            t = rng.uniform(0, 2 * math.pi)
            r = sh.radius_at(t)
            out.append(_line([(cx, cy), (cx + r * math.cos(t), cy + r * math.sin(t))], to_disp, rng.uniform(0.6, 1.4) * px_per_pt, colour(),
                             rng.uniform(0.8, 0.9), clip))
    elif style == 'grid':
        for k in range(n):
            f = (k + 1) / (n + 1) * 2 - 1
            if k % 2 == 0:      # horizontal chord at height f * r(pi/2)
                y = f * sh.radius_at(math.pi / 2) * 0.9
                xl, xr = -sh.radius_at(math.pi) * 0.7, sh.radius_at(0.0) * 0.7
                out.append(_line([(cx + xl, cy + y), (cx + xr, cy + y)], to_disp, rng.uniform(0.6, 1.4) * px_per_pt, colour(), rng.uniform(0.8, 0.9), clip))
            else:
                x = f * sh.radius_at(0.0) * 0.9
                yb, yt = -sh.radius_at(-math.pi / 2) * 0.7, sh.radius_at(math.pi / 2) * 0.7
                out.append(_line([(cx + x, cy + yb), (cx + x, cy + yt)], to_disp, rng.uniform(0.6, 1.4) * px_per_pt, colour(), rng.uniform(0.8, 0.9), clip))
    elif style == 'polygon':
        m = rng.randint(3, 8)
        t0 = rng.uniform(0, 2 * math.pi)
        pts = []
        for k in range(m):
            t = t0 + 2 * math.pi * k / m
            r = 0.55 * sh.radius_at(t)
            pts.append((cx + r * math.cos(t), cy + r * math.sin(t)))
        pts.append(pts[0])
        out.append(_line(pts, to_disp, rng.uniform(0.6, 1.4) * px_per_pt, colour(), rng.uniform(0.8, 0.9), clip))
    else:
        for k in range(n):
            ta, tb = rng.uniform(0, 2 * math.pi), rng.uniform(0, 2 * math.pi)
            ra, rb = sh.radius_at(ta) * 0.98, sh.radius_at(tb) * 0.98
            out.append(_line([(cx + ra * math.cos(ta), cy + ra * math.sin(ta)), (cx + rb * math.cos(tb), cy + rb * math.sin(tb))], to_disp,
                             rng.uniform(0.6, 1.4) * px_per_pt, colour(), rng.uniform(0.8, 0.9), clip))
    return out


def _masks(rng: random.Random, sh: _Shape, center, to_disp, lw_px, clip) -> List[Dict[str, Any]]:
    """Occluders: a white polygon fill over the outline plus the occluder's boundary segment inside the shape, in black."""
    out = []
    cx, cy = center
    for _ in range(rng.randint(1, 3)):
        t = rng.uniform(0, 2 * math.pi)
        r = sh.radius_at(t)
        ox, oy = cx + r * math.cos(t), cy + r * math.sin(t)
        ext = rng.uniform(0.5, 1.2) * 0.6
        if rng.random() < 0.5:
            m = 24
            poly = [(ox + ext * math.cos(2 * math.pi * k / m), oy + ext * math.sin(2 * math.pi * k / m)) for k in range(m)]
        else:
            poly = [(ox - ext, oy - ext * 0.7), (ox + ext, oy - ext * 0.7), (ox + ext, oy + ext * 0.7), (ox - ext, oy + ext * 0.7)]
        v, c = fp.polyline(poly, closed=True)
        out.append(dict(vertices=v, codes=c, affine=to_disp, linewidth=0.0, edgecolor=(0, 0, 0, 0), facecolor=(1.0, 1.0, 1.0, 1.0), clip=clip))
        # the part of the occluder's boundary that lies inside the shape
        inside = [p for p in poly + poly[:1] if math.hypot(p[0] - cx, p[1] - cy) <= sh.radius_at(math.atan2(p[1] - cy, p[0] - cx)) + 1e-9]
        if len(inside) >= 2:
            out.append(_line(inside, to_disp, lw_px, (0, 0, 0), 1.0, clip, cap='butt', join='miter'))
    return out


def single_shape_image(rng: random.Random, W: int, H: int) -> List[Dict[str, Any]]:
    dpi = W / 8.0
    px_per_pt = dpi / 72.0
    to_data = fp.data_to_display(W, H)
    clip = (float(np.dot(to_data, [-5, -5, 1])[0]), float(np.dot(to_data, [-5, -5, 1])[1]),
            float(np.dot(to_data, [5, 5, 1])[0]), float(np.dot(to_data, [5, 5, 1])[1]))
    sh = _base_shape(rng, 3.0 * GLOBAL_SCALE)
    lw_px = rng.uniform(1.5, 2.0) * px_per_pt
    aff = fp.compose(sh.patch_affine, to_data)
    artists = [dict(vertices=sh.path[0], codes=sh.path[1], affine=aff, linewidth=lw_px, cap='butt', join='miter',
                    edgecolor=(0.0, 0.0, 0.0, 0.9), facecolor=None, clip=clip)]
    if rng.random() < 0.7:
        artists += _decorations(rng, sh, (0.0, 0.0), to_data, px_per_pt, clip)
    else:
        artists += _masks(rng, sh, (0.0, 0.0), to_data, lw_px, clip)
    return artists


def multi_shape_image(rng: random.Random, W: int, H: int) -> List[Dict[str, Any]]:
    dpi = W / 8.0
    px_per_pt = dpi / 72.0
    to_data = fp.data_to_display(W, H)
    clip = (float(np.dot(to_data, [-5, -5, 1])[0]), float(np.dot(to_data, [-5, -5, 1])[1]),
            float(np.dot(to_data, [5, 5, 1])[0]), float(np.dot(to_data, [5, 5, 1])[1]))
    n = rng.randint(2, 3)
    mode = rng.choice(('nested', 'adjacent', 'intersecting'))
    lw_px = rng.uniform(1.5, 2.0) * px_per_pt
    sizes = [rng.uniform(2.2, 3.2) * GLOBAL_SCALE for _ in range(n)]
    centres: List[Tuple[float, float]] = []
    if mode == 'nested':
        sizes = [sizes[0] * (0.4 ** i) * (1.0 if i == 0 else 1.6) for i in range(n)]
        centres = [(0.0, 0.0)] * n
    elif mode == 'adjacent':
        total = sum(sizes)
        x = -total / 2
        for s in sizes:
            centres.append((x + s / 2, 0.0))
            x += s
        k = min(1.0, 9.0 / total)
        sizes = [s * k for s in sizes]
        centres = [(c[0] * k, c[1] * k) for c in centres]
    else:
        for i, s in enumerate(sizes):
            t = rng.uniform(0, 2 * math.pi)
            d = 0.0 if i == 0 else rng.uniform(0.3, 0.6) * s
            centres.append((d * math.cos(t), d * math.sin(t)))
    artists = []
    for s, c in zip(sizes, centres):
        sh = _base_shape(rng, s, c)
        aff = sh.patch_affine
        if rng.random() < 0.4:      # StyleEnhancer.rotate: rotate_deg_around the shape centre
            ang = rng.uniform(0, 360)
            aff = fp.compose(aff, fp.affine_translate(-c[0], -c[1]), fp.affine_rotate_deg(ang), fp.affine_translate(c[0], c[1]))
        artists.append(dict(vertices=sh.path[0], codes=sh.path[1], affine=fp.compose(aff, to_data), linewidth=lw_px, cap='butt', join='miter',
                            edgecolor=(0.0, 0.0, 0.0, 0.9), facecolor=None, clip=clip))
    return artists


def synthetic_figure_scene(n_images: int, W: int = 512, H: int = 512, seed: int = 0, kind: str = 'single', first: int = 0) -> FigureScene:
    """Images [first, first + n_images) of the synthetic scene `seed`; kind: 'single' (config 2) or 'multi' (config 3)."""
    sb = FigureSceneBuilder(W, H)
    make = single_shape_image if kind == 'single' else multi_shape_image
    for i in range(first, first + n_images):
        rng = random.Random(seed * 1000003 + i)
        sb.add_image(make(rng, W, H))
    return sb.finish()


def tile_scene(scene: FigureScene, times: int) -> FigureScene:
    """The scene repeated `times` times (bench.py builds large resident scenes from a few thousand distinct images)."""
    if times <= 1:
        return scene
    na, nv = scene.n_artists, scene.n_vertices
    iab = np.concatenate([scene.image_artist_begin[:-1].astype(np.int64) + k * na for k in range(times)] + [[times * na]])
    avb = np.concatenate([scene.artist_vertex_begin[:-1].astype(np.int64) + k * nv for k in range(times)] + [[times * nv]])
    rep = lambda a: np.concatenate([a] * times)   # noqa: E731
    return FigureScene(scene.W, scene.H, iab, avb, rep(scene.vertices), rep(scene.codes), rep(scene.affine), rep(scene.linewidth),
                       rep(scene.cap), rep(scene.join), rep(scene.snap), rep(scene.edge_rgba), rep(scene.face_rgba), rep(scene.has_face),
                       rep(scene.clip), rep(scene.has_clip))
