from .artist import Artist
from .axis import Axis
from .patches import Patch
from .spines import Spine
from .text import Text
from .transforms import Bbox


class _FigPatch:
    def get_visible(self): return True
    def get_facecolor(self): return (1.0, 1.0, 1.0, 1.0)


class Axes(Artist):
    def __init__(self, fig):
        super().__init__()
        self.figure, self.axison, self._children = fig, False, []
        self.patch = Patch(None, None, facecolor='white')
        self._fixed = [Spine(), Spine(), Axis(), Axis(), Text(''), Text(''), Text(''), self.patch]
        self.aspect_applied = 0

    def apply_aspect(self): self.aspect_applied += 1
    def add(self, artist):
        artist.figure = self.figure
        self._children.append(artist)
        return artist

    def get_children(self): return self._children + self._fixed


class Figure:
    def __init__(self, dpi=64.0, size=512):
        self.dpi, self.bbox, self.patch = dpi, Bbox(0, 0, size, size), _FigPatch()
        self.axes, self.artists, self.lines, self.patches, self.texts, self.images, self.legends = [], [], [], [], [], [], []

    def set_dpi(self, dpi): self.dpi = dpi
    def add_axes(self):
        ax = Axes(self)
        self.axes.append(ax)
        return ax
