"""ORACLE (test infrastructure, not product code) -- one-world differentiable stepping.

CPU float64 torch restatement of the reference's ``World3D.step`` hot path for a
SINGLE world (the reference has no other mode), differentiable by autograd so
it also serves as the gradient oracle.  Each function cites what it follows:

* bodies / integrator      sdf_physics/physics3d/bodies.py:398-511, 627-760
* SDF contact search       sdf_physics/physics3d/contacts.py:27-272
* contact Jacobians        sdf_physics/physics3d/world.py:48-101
* engine (LCP assembly)    lcp_physics/physics/engines.py:31-83
* step state machine, time-of-contact differential
                            lcp_physics/physics/world.py:119-139, 141-237, 241-379

Meshes and SDF grids are INPUTS: the caller hands in vertex /
face buffers; nothing here meshes an SDF.  Broad phase = AABB overlap of the
rotated cube of half side ``scale + eps``, pairs (i<j) in body-list order
(declared semantics of the py3ode stand-in, tests/golden/ref_shims/ode.py).

Pinned against the unmodified reference executed with third-party stand-ins:
tests/golden/make_golden.py -> tests/golden/*.npz, checked by
tests/test_oracle_golden.py.
"""
import numpy as np
import torch
from torch.nn.functional import normalize
from scipy.spatial import ConvexHull
try:                                    # scipy >= 1.8 exposes QhullError here
    from scipy.spatial import QhullError
except ImportError:                     # pragma: no cover
    from scipy.spatial.qhull import QhullError

from . import sdf as S
from . import transforms as T
from .lcp import make_lcp_function

DEFAULT_EPS = 1e-3     # Defaults3D.EPSILON  sdf_physics/physics3d/utils.py:46
DEFAULT_TOL = 1e-8     # Defaults3D.TOL      sdf_physics/physics3d/utils.py:49
F64 = torch.float64


def tens(x):
    return x if isinstance(x, torch.Tensor) else torch.tensor(x, dtype=F64)


class Body:
    """Pose p=[qw,qx,qy,qz,x,y,z], velocity v=[w,v]; SDF of ``kind`` with normalised params."""

    def __init__(self, kind, params, scale, verts, faces, pos, vel=(0, 0, 0, 0, 0, 0), mass=1.0,
                 ang_inertia=None, restitution=0.5, fric_coeff=0.9, eps=DEFAULT_EPS):
        self.kind, self.params, self.scale = kind, [tens(a) for a in params], tens(scale)
        self.verts, self.faces = tens(verts), torch.as_tensor(faces).long()
        pos = tens(pos)
        self.p = torch.cat([pos.new_tensor([1., 0, 0, 0]), pos]) if pos.numel() == 3 else pos
        vel = tens(vel)
        self.v = torch.cat([vel.new_zeros(3), vel]) if vel.numel() == 3 else vel
        self.mass = tens(mass)
        self.ang_inertia = tens(ang_inertia)
        self.restitution, self.fric_coeff = tens(restitution), tens(fric_coeff)
        self.eps = float(eps)
        self.forces = []            # callables t -> (6,) tensor  [torque; force]
        self.no_contact = set()
        self.M = None
        self.set_p(self.p)

    # views ------------------------------------------------------------
    @property
    def rot(self):
        return self.p[:4]

    @property
    def pos(self):
        return self.p[4:]

    def set_p(self, p):
        """bodies.py:498-511 (pose + world-frame inertia block)."""
        self.p = p
        R = T.quaternion_to_matrix(self.rot)
        self.M = torch.block_diag(R @ self.ang_inertia @ R.t(), torch.eye(3, dtype=F64) * self.mass)

    def move(self, dt):
        """bodies.py:488-496."""
        dq = T.matrix_to_quaternion(T.so3_exponential_map(self.v[:3].unsqueeze(0) * dt))
        self.set_p(torch.cat([T.quaternion_multiply(dq, self.p[:4]).squeeze(), self.p[4:] + self.v[3:] * dt]))

    def force(self, t):
        if not self.forces:
            return self.v.new_zeros(6)
        return sum(f(t) for f in self.forces)

    def add_gravity(self, g=10.0):
        """physics3d/forces.py:73-85: [0,0,0,0,-m g,0]."""
        self.forces.append(lambda t: tens([0., 0, 0, 0, -1, 0]) * self.mass * g)

    def add_no_contact(self, other):
        self.no_contact.add(other)
        other.no_contact.add(self)

    def world_verts(self):
        return T.quaternion_apply(self.rot, self.verts) + self.pos

    def query(self, pts, want_dir=True):
        return S.query(self.kind, self.params, self.scale, pts, want_dir)

    def aabb_half(self):
        R = T.quaternion_to_matrix(self.rot.detach())
        return R.abs() @ ((self.scale.detach() + self.eps) * torch.ones(3, dtype=F64))


# ------------------------------------------------------------------ contacts
def cubes_overlap(b1, b2):
    """contacts.py:27-36."""
    v1 = T.quaternion_apply(T.quaternion_invert(b2.rot), b1.world_verts() - b2.pos)
    v2 = T.quaternion_apply(T.quaternion_invert(b1.rot), b2.world_verts() - b1.pos)
    a = torch.any(torch.all((-b2.scale <= v1) & (v1 <= b2.scale), dim=1))
    b = torch.any(torch.all((-b1.scale <= v2) & (v2 <= b1.scale), dim=1))
    return bool(a and b)


def frank_wolfe(b1, b2, eps, tol, stats=None):
    """contacts.py:39-94 -> (barycentrics (K,3), face ids (K,))."""
    verts = T.quaternion_apply(T.quaternion_invert(b2.rot), b1.world_verts() - b2.pos)
    tri = verts[b1.faces]                                # (F,3,3)
    c = tri.mean(dim=1)
    c_sdf, c_dir = b2.query(c)
    rad = (c.unsqueeze(1) - tri).norm(dim=2).max(dim=1)[0]
    cand = (c_sdf < rad + eps) & (c_dir.norm(dim=1) > 1e-12)
    if stats is not None:
        stats['candidates'] = torch.nonzero(cand).squeeze(1)
    if not torch.any(cand):
        return verts.new_empty((0, 3)), torch.nonzero(cand, as_tuple=False).squeeze(1)
    pqr = tri[cand]
    K = pqr.shape[0]
    ar = torch.arange(K)
    corner = b2.query(pqr.reshape(-1, 3), want_dir=False).reshape(K, 3).argmin(dim=1)
    x = pqr[ar, corner]
    abc = x.new_zeros((K, 3))
    abc[ar, corner] = 1.
    n_it = 0
    for k in range(32):
        val, g = b2.query(x)
        pick = (pqr @ g.unsqueeze(2)).squeeze(2).argmin(dim=1)
        s = pqr[ar, pick]
        gain = ((x - s).unsqueeze(1) @ g.unsqueeze(2)).reshape(K)
        gamma = (2.0 / (k + 2.0)) * (gain.abs() > tol).to(F64)
        if torch.all(gamma == 0) or torch.any(val < -tol):
            break
        n_it += 1
        x = (1.0 - gamma).unsqueeze(1) * x + gamma.unsqueeze(1) * s
        abc = abc * (1.0 - gamma).unsqueeze(1)
        abc[ar, pick] += gamma
    if stats is not None:
        stats['fw_iters'] = n_it
    # b1 is always an SDF body here: push from the triangle onto b1's true surface
    x1 = (b1.verts[b1.faces[cand]] * abc.unsqueeze(2)).sum(dim=1)
    d1, g1 = b1.query(x1)
    rel = T.quaternion_multiply(T.quaternion_invert(b2.rot), b1.rot)
    x = x - d1.unsqueeze(1) * T.quaternion_apply(rel, g1)
    keep = b2.query(x, want_dir=False) <= eps
    cand = cand.clone()
    cand[cand.clone()] &= keep
    return abc[keep], torch.nonzero(cand, as_tuple=False).squeeze(1)


def contact_geometry(b1, b2, abc, ids, fd_eps=DEFAULT_EPS, detach_b2=False):
    """contacts.py:161-214 -> (normals, p1, p2, pen); differentiable w.r.t. poses/verts/params."""
    if ids.numel() == 0:
        e = abc.new_tensor([])
        return e, e, e, e
    c1 = (b1.verts[b1.faces[ids]] * abc.unsqueeze(2)).sum(dim=1)
    d1, n1 = b1.query(c1)
    c1 = c1 - d1.unsqueeze(1) * n1
    d1, n1 = b1.query(c1)
    cw = T.quaternion_apply(b1.rot, c1) + b1.pos
    c2 = T.quaternion_apply(T.quaternion_invert(b2.rot), cw - b2.pos)
    if detach_b2:
        c2 = c2.detach()
    d2, n2 = b2.query(c2)

    def laplacian(body, c, d0):
        acc = c.new_zeros(c.shape[0])
        for ax in range(3):
            sh = c.new_zeros(3)
            sh[ax] = fd_eps
            acc = acc + body.query(c + sh, want_dir=False) - 2 * d0 + body.query(c - sh, want_dir=False)
        return acc

    stable = laplacian(b2, c2, d2).abs() < laplacian(b1, c1, d1).abs()
    normals = T.quaternion_apply(b2.rot, n2) * stable.unsqueeze(1) \
        - T.quaternion_apply(b1.rot, n1) * (~stable).unsqueeze(1)
    p2 = T.quaternion_apply(b2.rot, c2 - d2.unsqueeze(1) * n2)
    p1 = T.quaternion_apply(b1.rot, c1)
    return normals, p1, p2, -d2


def hull_filter(normals, p1, eps):
    """contacts.py:97-158: cluster by normal (1e-2 rad), keep convex-hull vertices per cluster."""
    idx = torch.arange(normals.shape[0])
    if normals.shape[0] <= 1:
        return idx
    ok = normals.norm(dim=1) > 1e-12
    normals, p1, idx = normals[ok], p1[ok], idx[ok]
    kept = []
    while normals.shape[0] > 0:
        ang = torch.acos(torch.min(normals @ normals[0], normals.new_tensor(1.)))
        m = ang < 1e-2
        pts, ids = p1[m].detach(), idx[m]
        normals, p1, idx = normals[~m], p1[~m], idx[~m]
        while True:
            if pts.shape[1] > 1:
                try:
                    sel = torch.as_tensor(np.asarray(ConvexHull(pts.numpy()).vertices)).long()
                    break
                except (QhullError, ValueError):
                    drop = int(pts.var(dim=0).argmin())
                    pts = pts[:, [a for a in range(pts.shape[1]) if a != drop]]
            else:
                lo, hi = pts.argmin(), pts.argmax()
                sel = torch.stack([lo, hi]) if (pts.max() - pts.min()) > eps else torch.stack([lo])
                break
        kept.append(ids[sel])
    return torch.cat(kept) if kept else torch.empty(0, dtype=torch.long)


# --------------------------------------------------------------------- world
def tangent_seed(n):
    """physics3d/utils.py:247-256: e_k x n with k = argmin |n_k| (first on ties)."""
    k = int(n.abs().argmin())
    e = torch.zeros(3, dtype=n.dtype)
    e[k] = 1.
    return torch.linalg.cross(e, n)


class TimeOfContact(torch.autograd.Function):
    """world.py:141-237: identity on dt forward; implicit time-of-impact derivative backward."""

    @staticmethod
    def gap(h, c1, c2, v1, v2, x1, x2, R1, R2, n2, a1, a2):
        Rih = T.so3_exponential_map(h * v1[:, :3]) @ R1
        Rjh = T.so3_exponential_map(h * v2[:, :3]) @ R2
        xi = x1 + h * v1[:, 3:] + 0.5 * a1[:, 3:] * h * h
        xj = x2 + h * v2[:, 3:] + 0.5 * a2[:, 3:] * h * h
        ci_w = Rih @ c1.unsqueeze(2) + xi.unsqueeze(2)
        ci_j = (Rjh.transpose(1, 2) @ (ci_w - xj.unsqueeze(2))).squeeze(2)
        return (n2.unsqueeze(1) @ (c2 - ci_j).unsqueeze(2)).squeeze(2)

    @staticmethod
    def forward(ctx, h, *args):
        ctx.save_for_backward(h, *args)
        return h

    @staticmethod
    def backward(ctx, gh):
        saved = ctx.saved_tensors
        h = saved[0]
        with torch.enable_grad():
            jac = torch.autograd.functional.jacobian(TimeOfContact.gap, saved, strict=True)
        dD_dh = jac[0].clone()
        # world.py:204 reads Defaults.TOL of the 2-D base module (lcp_physics/physics/utils.py:43 = 1e-6),
        # not Defaults3D.TOL
        dD_dh[dD_dh < 1e-6 / h] = 0.
        den = torch.sum(dD_dh ** 2, dim=0)
        w = dD_dh / den if den > 1e-5 else 0. * dD_dh
        outs = [gh]
        for j in jac[1:]:
            ww = -w.reshape(w.shape + (1,) * (j.dim() - w.dim()))
            outs.append(torch.sum(ww * j, dim=0).squeeze(0) * gh)
        return tuple(outs)


class World:
    """World3D restated: same constructor knobs and step semantics."""

    def __init__(self, bodies, pinned=(), axis_locks=(), dt=1.0 / 30, eps=DEFAULT_EPS, tol=DEFAULT_TOL,
                 fric_dirs=8, strict_no_penetration=True, time_of_contact_diff=True,
                 stop_contact_grad=False, stop_friction_grad=False, detach_contact_b2=False, max_iter=10,
                 post_stab=False, grippers=()):
        self.bodies = list(bodies)
        self.post_stab = post_stab
        self.grippers = [(int(i1), int(i2), tens(list(axis))) for i1, i2, axis in grippers]
        self.nb = len(self.bodies)
        # equality rows: TotalConstraint3D (6 rows, J = I6) then single-axis locks (body, axis in 0..5)
        rows = []
        for b in pinned:
            i = self.bodies.index(b)
            for a in range(6):
                rows.append((i, a))
        for b, a in axis_locks:
            rows.append((self.bodies.index(b), a))
        self.eq_rows = rows
        self.t, self.dt, self.last_dt = 0, dt, None
        self.eps, self.tol, self.fric_dirs = eps, tol, fric_dirs
        self.strict = strict_no_penetration
        self.toc_diff = time_of_contact_diff
        self.stop_contact_grad, self.stop_friction_grad = stop_contact_grad, stop_friction_grad
        self.detach_contact_b2 = detach_contact_b2
        self.max_iter = max_iter
        self.v = torch.cat([b.v for b in self.bodies])
        self.contacts, self.toc_contacts = None, None
        self.trajectory = []
        self.stats = {'substeps': [], 'lcp_sizes': []}
        self.find_contacts()
        if self.strict:
            assert all(c[0][3].item() <= self.tol for c in self.contacts), 'Interpenetration at start'

    # state ------------------------------------------------------------
    def set_v(self, v):
        self.v = v
        for i, b in enumerate(self.bodies):
            b.v = v[6 * i:6 * i + 6]

    def get_p(self):
        return torch.cat([b.p for b in self.bodies])

    def set_p(self, p):
        for i, b in enumerate(self.bodies):
            b.set_p(p[7 * i:7 * i + 7])

    # contact detection ------------------------------------------------
    def _search(self, i1, i2):
        """contacts.py:248-272, direction mesh(b1) -> sdf(b2)."""
        b1, b2 = self.bodies[i1], self.bodies[i2]
        with torch.no_grad():
            abc, ids = frank_wolfe(b1, b2, self.eps, self.tol)
            self.last_prefilter.append((i1, i2, ids.clone()))
            n, p1, p2, pen = contact_geometry(b1, b2, abc, ids, detach_b2=self.detach_contact_b2)
            ok = bool(torch.all(pen <= self.tol))
            if ok:
                keep = hull_filter(n, p1, self.eps)
        if ok:
            n, p1, p2, pen = contact_geometry(b1, b2, abc[keep], ids[keep], detach_b2=self.detach_contact_b2)
            self.last_search.append((i1, i2, ids[keep].clone(), abc[keep].clone()))
        else:
            self.last_search.append((i1, i2, ids.clone(), abc.clone()))
        for k in range(len(pen)):
            self.contacts.append(((n[k], p1[k], p2[k], pen[k]), i1, i2))
        return ok

    def find_contacts(self):
        """world.py:396-399 + handler dispatch contacts.py:221-244."""
        self.contacts = []
        self.last_search = []
        self.last_prefilter = []
        half = [b.aabb_half() for b in self.bodies]
        for i in range(self.nb):
            for j in range(i + 1, self.nb):
                bi, bj = self.bodies[i], self.bodies[j]
                if not torch.all((bi.pos.detach() - bj.pos.detach()).abs() <= half[i] + half[j]):
                    continue
                if bi in bj.no_contact:
                    continue
                if not cubes_overlap(bi, bj):
                    continue
                if self._search(i, j):
                    self._search(j, i)

    # LCP ingredients ----------------------------------------------------
    def M(self):
        return torch.block_diag(*[b.M for b in self.bodies])

    def Je(self):
        J = torch.zeros(len(self.eq_rows), 6 * self.nb, dtype=F64)
        for r, (i, a) in enumerate(self.eq_rows):
            J[r, 6 * i + a] = 1.
        blocks = [J]
        for i1, i2, axis in self.grippers:
            # GripperJoint.J (physics3d/constraints.py:163-184): equal angular velocities; no relative linear motion of
            # the joint point (body1's origin) along the two directions orthogonal to the axis (body1 frame)
            b1, b2 = self.bodies[i1], self.bodies[i2]
            ax = T.quaternion_apply(b1.rot, axis)
            eye = torch.eye(3, dtype=F64)
            d1 = torch.linalg.cross(eye[int(ax.abs().argmin())], ax)
            d2 = torch.linalg.cross(d1, ax)
            dirs = normalize(torch.stack([d1, d2]), dim=1)
            pos2 = b1.pos - b2.pos
            z = pos2.new_zeros(())
            skew2 = torch.stack([torch.stack([z, -pos2[2], pos2[1]]), torch.stack([pos2[2], z, -pos2[0]]),
                                 torch.stack([-pos2[1], pos2[0], z])])
            Jg = torch.zeros(5, 6 * self.nb, dtype=F64)
            Jg[:3, 6 * i1:6 * i1 + 3] = eye
            Jg[:3, 6 * i2:6 * i2 + 3] = -eye
            Jg[3:, 6 * i1 + 3:6 * i1 + 6] = dirs
            Jg[3:, 6 * i2:6 * i2 + 3] = dirs @ skew2
            Jg[3:, 6 * i2 + 3:6 * i2 + 6] = -dirs
            blocks.append(Jg)
        return torch.cat(blocks) if len(blocks) > 1 else J

    def Jc(self):
        J = torch.zeros(len(self.contacts), 6 * self.nb, dtype=F64)
        for r, (c, i1, i2) in enumerate(self.contacts):
            n, p1, p2 = (a.detach() for a in c[:3]) if self.stop_contact_grad else c[:3]
            J[r, 6 * i1:6 * i1 + 6] = torch.cat([torch.linalg.cross(p1, n), n])
            J[r, 6 * i2:6 * i2 + 6] = -torch.cat([torch.linalg.cross(p2, n), n])
        return J

    def Jf(self):
        fd = self.fric_dirs
        J = torch.zeros(len(self.contacts) * fd, 6 * self.nb, dtype=F64)
        for r, (c, i1, i2) in enumerate(self.contacts):
            n, p1, p2 = (a.detach() for a in c[:3]) if self.stop_friction_grad else c[:3]
            d1 = normalize(tangent_seed(n), dim=0)
            d2 = normalize(torch.linalg.cross(d1, n), dim=0)
            dirs = torch.stack([d1, d2])
            if fd == 8:
                d3 = normalize(d1 + d2, dim=0)
                d4 = normalize(torch.linalg.cross(d3, n), dim=0)
                dirs = torch.cat([dirs, torch.stack([d3, d4])], dim=0)
            dirs = torch.cat([dirs, -dirs], dim=0)
            J[r * fd:(r + 1) * fd, 6 * i1:6 * i1 + 6] = torch.cat(
                [torch.linalg.cross(p1.expand(fd, -1), dirs), dirs], dim=1)
            J[r * fd:(r + 1) * fd, 6 * i2:6 * i2 + 6] = -torch.cat(
                [torch.linalg.cross(p2.expand(fd, -1), dirs), dirs], dim=1)
        return J

    def solve_dynamics(self, dt):
        """engines.py:31-83."""
        Je = self.Je()
        nz, neq = 6 * self.nb, Je.shape[0]
        f = torch.cat([b.force(self.t) for b in self.bodies])
        M = self.M()
        u = M @ self.v + dt * f
        if not self.contacts:
            if neq:
                P = torch.cat([torch.cat([M, -Je.t()], dim=1), torch.cat([Je, Je.new_zeros(neq, neq)], dim=1)])
                x = torch.inverse(P) @ torch.cat([u, u.new_zeros(neq)])
            else:
                x = torch.inverse(M) @ u
            self.last_lcp = None
            return x[:nz]
        nc, fd = len(self.contacts), self.fric_dirs
        Jc, Jf = self.Jc(), self.Jf()
        e = torch.stack([(self.bodies[i1].restitution + self.bodies[i2].restitution) / 2
                         for _, i1, i2 in self.contacts])
        mu = torch.stack([0.5 * (self.bodies[i1].fric_coeff + self.bodies[i2].fric_coeff)
                          for _, i1, i2 in self.contacts])
        G = torch.cat([Jc, Jf, Jf.new_zeros(nc, nz)])
        F = G.new_zeros(G.shape[0], G.shape[0])
        E = torch.zeros(fd * nc, nc, dtype=F64)
        for k in range(nc):
            E[k * fd:(k + 1) * fd, k] = 1
        F[nc:nc + fd * nc, nc + fd * nc:] = E
        F[nc + fd * nc:, :nc] = torch.diag(mu)
        F[nc + fd * nc:, nc:nc + fd * nc] = -E.t()
        h = torch.cat([(Jc @ self.v) * e, Jc.new_zeros(fd * nc + nc)])
        A = Je.unsqueeze(0) if neq else torch.tensor([])
        b = Je.new_zeros(1, neq) if neq else torch.tensor([])
        fn = make_lcp_function(max_iter=self.max_iter)
        self.stats['lcp_sizes'].append((nz, neq, G.shape[0]))
        z = fn(M.unsqueeze(0), u.unsqueeze(0), G.unsqueeze(0), h.unsqueeze(0), A, b, F.unsqueeze(0))
        self.last_lcp = dict(M=M, u=u, G=G, h=h, F=F, z=z[0])
        return -z[0]

    # stepping -----------------------------------------------------------
    def step(self, fixed_dt=True):
        """world.py:119-139."""
        self._undo = (self.get_p(), self.v, self.contacts, self.t)
        had = False
        n_sub = 0
        if fixed_dt:
            end_t = self.t + self.dt
            while self.t < end_t:
                n_sub += self.step_dt(end_t - self.t)
                had = had or bool(self.contacts)
        else:
            n_sub += self.step_dt(self.dt)
            had = bool(self.contacts)
        self.stats['substeps'].append(n_sub)
        return had

    def undo_step(self):
        p, v, c, t = self._undo
        self.t = t
        self.set_p(p.clone())
        self.set_v(v.clone())
        self.contacts = c
        while self.trajectory and self.trajectory[-1][0] > self.t:
            self.trajectory.pop()

    def step_dt(self, dt):
        """world.py:241-379; returns the number of solve attempts."""
        p0, v0, c0 = self.get_p(), self.v, self.contacts
        tries = 0
        while True:
            tries += 1
            dt_ = dt
            if self.toc_diff and self.toc_contacts:
                dt_ = -self.last_dt + (self.last_dt.detach() + dt_)
            self.set_v(self.solve_dynamics(dt_))
            for b in self.bodies:
                b.move(dt_)
            self.find_contacts()
            if all(c[0][3].item() <= self.tol for c in self.contacts):
                old_pairs = [{c[1], c[2]} for c in c0]
                self.toc_contacts = [c for c in self.contacts if {c[1], c[2]} not in old_pairs]
                if self.toc_diff and self.toc_contacts:
                    dt_ = self._time_of_contact(dt_, p0)
                break
            if not self.strict and dt < self.dt / 2 ** 10:
                break
            dt /= 2
            self.set_p(p0.clone())
            self.set_v(v0.clone())
            self.contacts = c0
        if self.post_stab:
            # world.py:358-370: half of the stabilising displacement, applied as a velocity over dt; contacts re-detected
            tmp_v = self.v
            self.set_v(self.post_stabilization() / 2)
            for b in self.bodies:
                b.move(dt)
            self.set_v(tmp_v)
            self.find_contacts()
        self.trajectory.append((self.t, self.get_p(), self.v, self.contacts))
        self.t += dt
        return tries

    def post_stabilization(self):
        """engines.py:85-121: min 1/2 z'Mz  s.t.  Jc z <= Jc v (1 - e),  Je z = Je v;  returns dp = -z (the reference builds
        a fresh LCPFunction() here: 20 iterations, not the engine's max_iter)."""
        v, M, Je = self.v, self.M(), self.Je()
        nz, neq = 6 * self.nb, Je.shape[0]
        ge = Je @ v
        if not self.contacts:
            if neq:
                P = torch.cat([torch.cat([M, -Je.t()], dim=1), torch.cat([Je, Je.new_zeros(neq, neq)], dim=1)])
                x = torch.inverse(P) @ torch.cat([Je.new_zeros(nz), ge])
            else:
                x = torch.inverse(M) @ M.new_zeros(nz)
            return -x[:nz]
        Jc = self.Jc()
        e = torch.stack([(self.bodies[i1].restitution + self.bodies[i2].restitution) / 2
                         for _, i1, i2 in self.contacts])
        gc = Jc @ v + (Jc @ v) * -e
        nc = Jc.shape[0]
        A = Je.unsqueeze(0) if neq else torch.tensor([])
        b = ge.unsqueeze(0) if neq else torch.tensor([])
        fn = make_lcp_function()
        z = fn(M.unsqueeze(0), M.new_zeros(1, nz), Jc.unsqueeze(0), gc.unsqueeze(0), A, b, Jc.new_zeros(1, nc, nc))
        return -z[0]

    def _time_of_contact(self, dt_, p0):
        """world.py:275-341."""
        tc = self.toc_contacts
        B = self.bodies
        v1 = torch.stack([B[c[1]].v for c in tc])
        v2 = torch.stack([B[c[2]].v for c in tc])
        c1 = torch.stack([c[0][1] for c in tc])
        c2 = torch.stack([c[0][2] for c in tc])
        x1 = torch.stack([B[c[1]].pos for c in tc]) - dt_ * v1[:, 3:]
        x2 = torch.stack([B[c[2]].pos for c in tc]) - dt_ * v2[:, 3:]
        a1 = torch.stack([B[c[1]].force(self.t) / B[c[1]].mass for c in tc])
        a2 = torch.stack([B[c[2]].force(self.t) / B[c[2]].mass for c in tc])
        n = torch.stack([c[0][0] for c in tc])
        R1 = T.so3_exponential_map(-dt_ * v1[:, :3]) @ T.quaternion_to_matrix(torch.stack([B[c[1]].rot for c in tc]))
        R2 = T.so3_exponential_map(-dt_ * v2[:, :3]) @ T.quaternion_to_matrix(torch.stack([B[c[2]].rot for c in tc]))
        c1 = (R1.transpose(1, 2) @ c1.unsqueeze(2)).squeeze(2)
        c2 = (R2.transpose(1, 2) @ c2.unsqueeze(2)).squeeze(2)
        n2 = (R2.transpose(1, 2) @ n.unsqueeze(2)).squeeze(2)
        if not torch.is_tensor(dt_):
            dt_ = c1.new_tensor(dt_)
        dt_ = TimeOfContact.apply(dt_, c1, c2, v1, v2, x1, x2, R1, R2, n2, a1, a2)
        self.set_p(p0.clone())
        for b in B:
            b.move(dt_)
        self.last_dt = dt_
        return dt_
