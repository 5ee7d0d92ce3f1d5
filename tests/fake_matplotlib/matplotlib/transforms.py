import numpy as np


class Affine2D:
    def __init__(self, matrix=None):
        self._m = np.eye(3) if matrix is None else np.asarray(matrix, dtype=float)

    def get_matrix(self):
        return self._m

    def get_affine(self):
        return self

    def transform_path_non_affine(self, path):
        return path

    def scale(self, sx, sy):
        return Affine2D(np.diag([sx, sy, 1.0]) @ self._m)

    def translate(self, tx, ty):
        t = np.eye(3)
        t[0, 2], t[1, 2] = tx, ty
        return Affine2D(t @ self._m)


class Bbox:
    def __init__(self, x0, y0, x1, y1):
        self.x0, self.y0, self.x1, self.y1 = x0, y0, x1, y1
        self.height = y1 - y0
