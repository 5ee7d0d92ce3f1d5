from .artist import Artist


class Axis(Artist):
    pass
