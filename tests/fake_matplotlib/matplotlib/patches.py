from .artist import Artist
from .colors import to_rgba


class Patch(Artist):
    zorder = 1

    def __init__(self, path, transform, facecolor=None, edgecolor='black', linewidth=1.0, alpha=None, capstyle='butt', joinstyle='miter',
                 linestyle='solid', hatch=None, fill=True):
        super().__init__()
        self._path, self._transform = path, transform
        self._alpha = alpha
        self._facecolor = to_rgba(facecolor, alpha) if facecolor not in (None, 'none') else (0.0, 0.0, 0.0, 0.0)
        self._edgecolor = None if edgecolor is None else (to_rgba(edgecolor, alpha) if edgecolor != 'none' else (0.0, 0.0, 0.0, 0.0))
        self._lw, self._cap, self._join, self._ls, self._hatch, self._fill = linewidth, capstyle, joinstyle, linestyle, hatch, fill

    def get_path(self): return self._path
    def get_transform(self): return self._transform
    def get_linewidth(self): return self._lw
    def get_edgecolor(self): return self._edgecolor if self._edgecolor is not None else (0.0, 0.0, 0.0, 0.0)
    def get_facecolor(self): return self._facecolor
    def get_fill(self): return self._fill
    def get_linestyle(self): return self._ls
    def get_capstyle(self): return self._cap
    def get_joinstyle(self): return self._join
    def get_hatch(self): return self._hatch
