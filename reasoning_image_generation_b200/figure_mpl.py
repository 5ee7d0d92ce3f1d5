"""figure_mpl.py -- matplotlib-side walker: a live ``Figure`` -> the artist list `FigureSceneBuilder` takes.

This is the drop-in seam of the multigraph path: ``GeometryGenerator._save_figure(fig, save_path, dpi)``
(/root/reference/multigraph_generation/generator.py:250-259) calls ``fig.savefig``; `extract_figure(fig, dpi)` reads from
the same figure exactly what ``RendererAgg.draw_path`` would be handed for each patch and line -- path, transform split into
its non-affine and affine parts, line width in pixels, cap / join / snap, RGBA colours with the artist alpha folded in, clip
rectangle -- in matplotlib's draw order (zorder, then insertion).  `savefig_batch` replaces the per-figure savefig loop.

matplotlib is NOT installed in this image, so this module cannot be exercised against the real library here: it is
written against matplotlib's documented artist API and is gated -- importing it without matplotlib raises ImportError, nothing
else in the package depends on it, tests/test_figure_mpl.py skips itself unless matplotlib is importable, and
tests/test_figure_mpl_walker.py runs its control flow (draw order, what is read, what is refused) against a stand-in.
Reference bug B1 (`math.lg10`, multi_combinator.py:127,1087) makes the reference's `adjacent` mode raise before it reaches
savefig; that is host logic the walker never sees (a figure that was never completed is never extracted).
"""
from __future__ import annotations

from typing import Any, Dict, List

import numpy as np

import matplotlib  # noqa: F401  (ImportError here is the gate)
from matplotlib.colors import to_rgba
from matplotlib.lines import Line2D
from matplotlib.patches import Patch


def _clip(artist, fig_h):
    box = artist.get_clip_box() if artist.get_clip_on() else None
    if box is None:
        return None
    return (float(box.x0), float(box.y0), float(box.x1), float(box.y1))


def _snap(artist):
    s = artist.get_snap()
    return None if s is None else bool(s)


def _patch_artist(p: Patch) -> Dict[str, Any]:
    # Patch.draw: path = self.get_path(); transform = self.get_transform(); tpath = transform.transform_path_non_affine(path);
    # affine = transform.get_affine(); renderer.draw_path(gc, tpath, affine, rgbFace)
    transform = p.get_transform()
    tpath = transform.transform_path_non_affine(p.get_path())
    affine = transform.get_affine().get_matrix()
    lw_px = p.get_linewidth() * p.figure.dpi / 72.0
    edge = p.get_edgecolor()           # already carries the artist alpha (Patch._set_edgecolor)
    # Patch._draw_paths_with_artist_properties: lw = 0 when the edge colour is fully transparent or the line style is 'None'
    if p._edgecolor is None or edge[3] == 0 or lw_px == 0 or p.get_linestyle() in ('None', 'none', ' ', ''):
        lw_px = 0.0
    face = p.get_facecolor() if p.get_fill() and p.get_facecolor()[3] != 0 else None
    codes = tpath.codes
    return dict(vertices=np.asarray(tpath.vertices, dtype=np.float64), codes=None if codes is None else np.asarray(codes, np.uint8),
                affine=affine, linewidth=float(lw_px), cap=p.get_capstyle(), join=p.get_joinstyle(), snap=_snap(p),
                edgecolor=tuple(edge), facecolor=None if face is None else tuple(face), clip=_clip(p, p.figure.bbox.height))


def _line_artist(l: Line2D) -> Dict[str, Any]:
    # Line2D.draw (solid lines, no markers): tpath, affine = self._get_transformed_path().get_transformed_path_and_affine()
    l.recache_always() if hasattr(l, 'recache_always') else l.recache()
    tpath, affine = l._get_transformed_path().get_transformed_path_and_affine()
    rgba = to_rgba(l.get_color(), l.get_alpha())
    lw_px = l.get_linewidth() * l.figure.dpi / 72.0
    codes = tpath.codes
    return dict(vertices=np.asarray(tpath.vertices, dtype=np.float64), codes=None if codes is None else np.asarray(codes, np.uint8),
                affine=affine.get_matrix(), linewidth=float(lw_px), cap=l.get_solid_capstyle(), join=l.get_solid_joinstyle(),
                snap=_snap(l), edgecolor=rgba, facecolor=None, clip=_clip(l, l.figure.bbox.height))


def _drawn_children(ax):
    """The artists Axes.draw would draw, in its order: children in insertion order, stably sorted by zorder.  With
    ``ax.axis('off')`` (the reference: multigraph_generation/generator.py:490) the axes background, the spines, the two Axis
    objects and the (empty) title texts are not drawn; anything else this walker cannot express raises."""
    from matplotlib.axis import Axis
    from matplotlib.spines import Spine
    from matplotlib.text import Text
    if ax.axison:
        raise NotImplementedError('axes frame / ticks / background are not extracted (the reference calls ax.axis("off"))')
    out = []
    for a in ax.get_children():
        if not a.get_visible():
            continue
        if a is ax.patch or isinstance(a, (Spine, Axis)):
            continue                                           # not drawn when the axes are off
        if isinstance(a, Text):
            if a.get_text() == '':
                continue                                       # the three title slots
            raise NotImplementedError('text artists are not extracted')
        if not isinstance(a, (Line2D, Patch)):
            raise NotImplementedError(f'{type(a).__name__} artists are not extracted (collections, images, legends, ...)')
        cp = a.get_clip_path()
        if cp is not None:
            raise NotImplementedError('clip paths other than the axes rectangle are not extracted')
        out.append(a)
    out.sort(key=lambda a: a.get_zorder())                     # stable: insertion order within a zorder
    return out


def extract_figure(fig, dpi: float | None = None) -> List[Dict[str, Any]]:
    """The draw_path argument lists of every visible patch and solid line of `fig`, in draw order.  Text, images, collections,
    markers, dashed lines, hatches and non-rectangular clip paths are outside what the reference's generator produces and raise
    NotImplementedError."""
    if dpi is not None:
        fig.set_dpi(dpi)
    if fig.patch.get_visible():
        fc = fig.patch.get_facecolor()
        if tuple(round(float(c), 6) for c in fc) != (1.0, 1.0, 1.0, 1.0):
            raise NotImplementedError('figure backgrounds other than opaque white (savefig\'s default facecolor)')
    for a in list(fig.artists) + list(fig.lines) + list(fig.patches) + list(fig.texts) + list(fig.images) + list(fig.legends):
        if a.get_visible():
            raise NotImplementedError('figure-level artists are not extracted')
    artists: List[Dict[str, Any]] = []
    for ax in fig.axes:
        ax.apply_aspect()
        for a in _drawn_children(ax):
            if isinstance(a, Line2D):
                if a.get_linestyle() not in ('-', 'solid') or a.get_marker() not in (None, 'None', '', ' '):
                    raise NotImplementedError('dashed lines / markers')
                artists.append(_line_artist(a))
            else:
                if a.get_hatch():
                    raise NotImplementedError('hatches')
                artists.append(_patch_artist(a))
    return artists


def savefig_batch(figs, paths, renderer, encoder, dpi: float):
    """``for fig, path in zip(figs, paths): fig.savefig(path, dpi=dpi)`` on the GPU: one scene, one render + encode pass."""
    from .figure import FigureSceneBuilder
    from .png import write_files
    sb = FigureSceneBuilder(renderer.W, renderer.H)
    for fig in figs:
        sb.add_image(extract_figure(fig, dpi))
    scene = sb.finish()
    for i0, buf, off in renderer.iter_png(scene, encoder):
        write_files(buf, off, paths[i0:i0 + len(off) - 1])
