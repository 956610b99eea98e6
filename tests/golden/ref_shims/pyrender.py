class MetallicRoughnessMaterial:
    def __init__(self, **kw):
        pass


class Mesh:
    def __init__(self, tm):
        self._tm = tm

    @property
    def bounds(self):
        return self._tm.bounds

    @staticmethod
    def from_trimesh(tm, **kw):
        return Mesh(tm)
