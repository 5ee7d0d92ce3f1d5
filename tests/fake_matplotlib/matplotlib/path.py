import numpy as np


class Path:
    STOP, MOVETO, LINETO, CURVE3, CURVE4, CLOSEPOLY = 0, 1, 2, 3, 4, 79

    def __init__(self, vertices, codes=None):
        self.vertices = np.asarray(vertices, dtype=float).reshape(-1, 2)
        self.codes = None if codes is None else np.asarray(codes, dtype=np.uint8)
