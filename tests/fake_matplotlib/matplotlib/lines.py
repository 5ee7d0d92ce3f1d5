from .artist import Artist


class _TransformedPath:
    def __init__(self, path, transform):
        self._p, self._t = path, transform

    def get_transformed_path_and_affine(self):
        return self._p, self._t


class Line2D(Artist):
    zorder = 2

    def __init__(self, path, transform, color='C0', linewidth=1.5, alpha=None, linestyle='-', marker='None', solid_capstyle='projecting',
                 solid_joinstyle='round'):
        super().__init__()
        self._path, self._transform = path, transform
        self._color, self._lw, self._alpha, self._ls, self._marker, self._cap, self._join = color, linewidth, alpha, linestyle, marker, solid_capstyle, solid_joinstyle
        self.recached = 0

    def recache(self): self.recached += 1
    def _get_transformed_path(self): return _TransformedPath(self._path, self._transform)
    def get_color(self): return self._color
    def get_alpha(self): return self._alpha
    def get_linewidth(self): return self._lw
    def get_linestyle(self): return self._ls
    def get_marker(self): return self._marker
    def get_solid_capstyle(self): return self._cap
    def get_solid_joinstyle(self): return self._join
