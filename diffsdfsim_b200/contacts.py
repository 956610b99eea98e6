"""SDF collision query host side: geometry table, detection launch, differentiable contact geometry.

Mirrors the reference's contact-handler plug-in (sdf_physics/physics3d/contacts.py:217-272): the class
``FWContactHandler`` (alias ``B200ContactHandler``) is what ``World3D(contact_callback=...)`` resolves by name;
in this batched implementation it detects the contacts of ALL worlds with one launch sequence instead of
being called back per body pair by py3ode.
"""
import numpy as np
import torch

from . import _lib

F64 = torch.float64
_GEOM_DTYPE = np.dtype([('kind', 'i4'), ('nverts', 'i4'), ('nfaces', 'i4'), ('res', 'i4'), ('verts', 'u8'),
                        ('faces', 'u8'), ('grid', 'u8'), ('vstride', 'i8'), ('gstride', 'i8')], align=True)
assert _GEOM_DTYPE.itemsize == 56


class GeometryTable:
    """Device array of ``dsdf_body_geom`` (include/dsdf_b200.h) + the tensors it points into."""

    def __init__(self, bodies, W, device):
        self.keep = []
        rows = np.zeros(len(bodies), dtype=_GEOM_DTYPE)
        self.nfaces = []
        for i, b in enumerate(bodies):
            verts = b.verts.to(device=device, dtype=F64).contiguous()
            faces = b.faces.to(device=device, dtype=torch.int32).contiguous()
            per_world = verts.dim() == 3
            if per_world:
                assert verts.shape[0] == W
            nverts = verts.shape[-2]
            grid = getattr(b, 'sdf_grid', None)
            res, gptr, gstride = 0, 0, 0
            if grid is not None:
                grid = grid.to(device=device, dtype=F64).contiguous()
                res = grid.shape[-1]
                gstride = res ** 3 if grid.dim() == 4 else 0
                gptr = grid.data_ptr()
                self.keep.append(grid)
            self.keep += [verts, faces]
            rows[i] = (b.kind, nverts, faces.shape[0], res, verts.data_ptr(), faces.data_ptr(), gptr,
                       nverts * 3 if per_world else 0, gstride)
            self.nfaces.append(int(faces.shape[0]))
        self.rows = rows
        self.dev = torch.from_numpy(rows.view(np.uint8).copy()).to(device)

    def ptr(self):
        return self.dev.data_ptr()


class ContactSet:
    """Contacts of all worlds after one detection pass (padded to ``maxc`` per world)."""
    __slots__ = ('count', 'body', 'face', 'abc', 'geo', 'status', 'pre_ids', 'pre_cnt')

    def clone(self):
        c = ContactSet()
        for k in self.__slots__:
            v = getattr(self, k, None)
            setattr(c, k, v.clone() if isinstance(v, torch.Tensor) else v)
        return c


class ContactDetector:
    """Owns the work buffers of ``dsdf_contacts_detect`` for one batched world."""

    def __init__(self, table, pairs, W, nb, device, capK=768, maxc=16, record_prefilter=False):
        L = _lib.lib()
        self.table, self.W, self.nb, self.capK, self.maxc = table, W, nb, capK, maxc
        self.npairs = len(pairs)
        self.pairs = torch.tensor(pairs, dtype=torch.int32, device=device).reshape(-1, 2).contiguous()
        prefix = [0]
        for (i, j) in pairs:
            prefix.append(prefix[-1] + L.dsdf_contact_chunks_per_face_count(table.nfaces[i]))   # direction i -> j
            prefix.append(prefix[-1] + L.dsdf_contact_chunks_per_face_count(table.nfaces[j]))   # direction j -> i
        self.total_chunks = prefix[-1]
        self.prefix = torch.tensor(prefix, dtype=torch.int32, device=device)
        self.ws = torch.empty(L.dsdf_contacts_workspace_bytes(W, max(self.npairs, 1), capK) // 4 + 4,
                              dtype=torch.int32, device=device)
        self.device = device
        self.record_prefilter = record_prefilter

    def new_set(self):
        c, W, m, d = ContactSet(), self.W, self.maxc, self.device
        c.count = torch.zeros(W, dtype=torch.int32, device=d)
        c.body = torch.zeros(W, m, 2, dtype=torch.int32, device=d)
        c.face = torch.zeros(W, m, dtype=torch.int32, device=d)
        c.abc = torch.zeros(W, m, 3, dtype=F64, device=d)
        c.geo = torch.zeros(W, m, 10, dtype=F64, device=d)
        c.status = torch.zeros(W, dtype=torch.int32, device=d)
        if self.record_prefilter:
            c.pre_ids = torch.zeros(W, max(2 * self.npairs, 1), self.capK, dtype=torch.int32, device=d)
            c.pre_cnt = torch.zeros(W, max(2 * self.npairs, 1), dtype=torch.int32, device=d)
        else:
            c.pre_ids = c.pre_cnt = None
        return c

    def detect(self, p, shape, out, active=None, eps=1e-3, tol=1e-8, fd_eps=1e-3, body_eps=1e-3, detach_b2=False):
        """Fill ``out`` (a ContactSet) for the active worlds; values only (no autograd graph)."""
        L = _lib.lib()
        _lib.require_cuda(p, shape)
        rc = L.dsdf_contacts_detect(self.table.ptr(), _lib.ptr(self.pairs), _lib.ptr(self.prefix), self.total_chunks,
                                    self.npairs, _lib.ptr(p), _lib.ptr(shape), _lib.ptr(active), self.W, self.nb,
                                    eps, tol, fd_eps, body_eps, int(detach_b2), self.capK, self.maxc,
                                    _lib.ptr(out.count), _lib.ptr(out.body), _lib.ptr(out.face), _lib.ptr(out.abc),
                                    _lib.ptr(out.geo), _lib.ptr(out.status), _lib.ptr(out.pre_ids),
                                    _lib.ptr(out.pre_cnt), _lib.ptr(self.ws), _lib.stream())
        _lib.check(rc, 'dsdf_contacts_detect')
        return out


class _ContactGeometry(torch.autograd.Function):
    """Attach the detected contact geometry to the autograd graph of the poses.

    Forward returns the values the detection pass already computed (identical arithmetic to the reference's
    second, grad-enabled ``_compute_contacts``, contacts.py:262-264); backward is its VJP w.r.t. the poses.
    """

    @staticmethod
    def forward(ctx, p, shape, geo, count, body, face, abc, table, fd_eps, detach_b2):
        ctx.save_for_backward(p, shape, count, body, face, abc)
        ctx.table, ctx.fd_eps, ctx.detach_b2 = table, fd_eps, detach_b2
        return geo.clone()

    @staticmethod
    def backward(ctx, ggeo):
        L = _lib.lib()
        p, shape, count, body, face, abc = ctx.saved_tensors
        W, nb = p.shape[0], p.shape[1]
        gp = torch.empty_like(p)
        rc = L.dsdf_contact_geometry_backward(ctx.table.ptr(), _lib.ptr(p), _lib.ptr(shape), W, nb, ctx.fd_eps,
                                              int(ctx.detach_b2), geo_cap(body), _lib.ptr(count), _lib.ptr(body),
                                              _lib.ptr(face), _lib.ptr(abc), _lib.ptr(ggeo.contiguous()), _lib.ptr(gp),
                                              _lib.stream())
        _lib.check(rc, 'dsdf_contact_geometry_backward')
        return (gp,) + (None,) * 9


def geo_cap(body):
    return body.shape[1]


def differentiable_geometry(p, shape, cs, table, fd_eps=1e-3, detach_b2=False):
    """(W,maxc,10) contact geometry [n,p1,p2,pen] connected to ``p`` for autograd."""
    return _ContactGeometry.apply(p, shape, cs.geo, cs.count, cs.body, cs.face, cs.abc, table, fd_eps, detach_b2)


class FWContactHandler:
    """Name-compatible with sdf_physics.physics3d.contacts.FWContactHandler; batched worlds call ``detector``."""

    def __call__(self, args, geom1, geom2):
        raise RuntimeError('the batched World3D detects all contacts in one pass; per-pair callbacks are not used')


B200ContactHandler = FWContactHandler
