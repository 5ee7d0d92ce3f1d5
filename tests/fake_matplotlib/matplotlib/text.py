from .artist import Artist


class Text(Artist):
    zorder = 3

    def __init__(self, text=''):
        super().__init__()
        self._text = text

    def get_text(self): return self._text
