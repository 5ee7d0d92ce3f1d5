from .artist import Artist


class Spine(Artist):
    pass
