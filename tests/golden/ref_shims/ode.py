"""py3ode stand-in: broad-phase only.  AABB of the rotated GeomBox, pairs i<j in insertion order."""
import numpy as np


def _np(x):
    return np.asarray(x.detach().cpu().numpy() if hasattr(x, "detach") else x, dtype=np.float64)


class GeomBox:
    def __init__(self, space, lengths):
        self.lengths = _np(lengths).reshape(3)
        self.pos = np.zeros(3)
        self.quat = np.array([1.0, 0.0, 0.0, 0.0])
        self.body = None
        self.no_contact = set()

    def setPosition(self, p):
        self.pos = _np(p).reshape(3)

    def getPosition(self):
        return tuple(self.pos)

    def setQuaternion(self, q):
        self.quat = _np(q).reshape(4)

    def aabb_half(self):
        r, i, j, k = self.quat
        s = 2.0 / float(self.quat @ self.quat)
        R = np.array([[1 - s * (j * j + k * k), s * (i * j - k * r), s * (i * k + j * r)],
                      [s * (i * j + k * r), 1 - s * (i * i + k * k), s * (j * k - i * r)],
                      [s * (i * k - j * r), s * (j * k + i * r), 1 - s * (i * i + j * j)]])
        return np.abs(R) @ (0.5 * self.lengths)


class HashSpace:
    def __init__(self):
        self.geoms = []

    def add(self, g):
        self.geoms.append(g)

    def collide(self, arg, callback):
        n = len(self.geoms)
        for a in range(n):
            for b in range(a + 1, n):
                ga, gb = self.geoms[a], self.geoms[b]
                if np.all(np.abs(ga.pos - gb.pos) <= ga.aabb_half() + gb.aabb_half()):
                    callback(arg, ga, gb)
