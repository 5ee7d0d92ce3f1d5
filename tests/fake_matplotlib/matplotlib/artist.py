class Artist:
    zorder = 0

    def __init__(self):
        self.figure = None
        self._visible, self._snap, self._clip_on, self._clip_box, self._clip_path = True, None, True, None, None
        self._zorder = self.zorder

    def get_visible(self): return self._visible
    def set_visible(self, v): self._visible = v
    def get_zorder(self): return self._zorder
    def set_zorder(self, z): self._zorder = z
    def get_snap(self): return self._snap
    def set_snap(self, s): self._snap = s
    def get_clip_on(self): return self._clip_on
    def get_clip_box(self): return self._clip_box
    def get_clip_path(self): return self._clip_path
