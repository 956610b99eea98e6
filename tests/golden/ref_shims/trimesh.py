import numpy as np


class Trimesh:
    def __init__(self, vertices=None, faces=None, **kw):
        self.vertices = np.asarray(vertices, dtype=np.float64)
        self.faces = np.asarray(faces)

    @property
    def bounds(self):
        return np.stack([self.vertices.min(0), self.vertices.max(0)])
