"""Bodies with the reference's constructor surface (sdf_physics/physics3d/bodies.py:398-1009), batched over worlds.

Every per-body tensor carries ONE leading batch dim B in {1, W}: p (B,7) = [qw,qx,qy,qz,x,y,z], v (B,6) = [w,v],
mass (B,), ang_inertia (B,3,3), fric_coeff (B,), restitution (B,), scale (B,), shape (B,3).  ``World3D`` broadcasts
B = 1 bodies to its world count.  The surface mesh (verts (V,3) or (B,V,3), faces (F,3)) and, for grid bodies, the
SDF grid ((R,R,R) or (B,R,R,R)) are inputs: pass ``mesh=(verts, faces)`` or let the primitive generate the
``custom_mesh`` lattice of the reference's density (meshes.py).
"""
import math

import numpy as np
import torch

from . import meshes, ops
from .utils import Defaults3D, as_batched

F64 = torch.float64
BOX, SPHERE, CYLINDER, GRID, BOX_ROUNDED, BRICK, BOWL = 0, 1, 2, 3, 4, 5, 6


def euler_to_quat(e):
    """physics3d/utils.py:211-226 (quat(vec, 'wxyz'))."""
    phi, the, psi = [0.5 * float(a) for a in e]
    w = math.cos(phi) * math.cos(the) * math.cos(psi) + math.sin(phi) * math.sin(the) * math.sin(psi)
    x = math.sin(phi) * math.cos(the) * math.cos(psi) - math.cos(phi) * math.sin(the) * math.sin(psi)
    y = math.cos(phi) * math.sin(the) * math.cos(psi) + math.sin(phi) * math.cos(the) * math.sin(psi)
    z = math.cos(phi) * math.cos(the) * math.sin(psi) - math.sin(phi) * math.sin(the) * math.cos(psi)
    return [w, x, y, z]


class Body3D:
    """bodies.py:398-511 (pose/velocity/mass bookkeeping; stepping itself happens in World3D's kernels)."""

    def __init__(self, pos, vel=(0, 0, 0, 0, 0, 0), mass=1, restitution=Defaults3D.RESTITUTION,
                 fric_coeff=Defaults3D.FRIC_COEFF, eps=Defaults3D.EPSILON, device=None):
        self.device = device
        pos = as_batched(pos, 1, device)
        n = pos.shape[1]
        if n == 3:
            q = pos.new_zeros(pos.shape[0], 4)
            q[:, 0] = 1
            self.p = torch.cat([q, pos], 1)
        elif n == 6:
            q = torch.stack([pos.new_tensor(euler_to_quat(r[:3].tolist())) for r in pos])
            self.p = torch.cat([q, pos[:, 3:]], 1)
        else:
            assert n == 7
            self.p = pos
        vel = as_batched(vel, 1, device)
        self.v = torch.cat([vel.new_zeros(vel.shape[0], 3), vel], 1) if vel.shape[1] == 3 else vel
        self.mass = as_batched(mass, 0, device)
        self.fric_coeff = as_batched(fric_coeff, 0, device)
        self.restitution = as_batched(restitution, 0, device)
        self.eps = float(eps)
        self.forces = []
        self.no_contact = set()
        self.ang_inertia = self._get_ang_inertia(self.mass)
        self._world = None
        self._index = None

    # views (reference attribute names) -----------------------------------------------------------
    @property
    def rot(self):
        return self.p[..., :4]

    @property
    def pos(self):
        return self.p[..., 4:]

    def set_p(self, new_p):
        self.p = as_batched(new_p, 1, self.p.device)
        world = self._world() if self._world is not None else None     # weak reference: no body <-> world cycle
        if world is not None:
            world._body_pose_changed(self._index)

    def add_force(self, f):
        self.forces.append(f)
        f.set_body(self)

    def add_no_contact(self, other):
        self.no_contact.add(other)
        other.no_contact.add(self)

    def apply_forces(self, t):
        """lcp_physics/physics/bodies.py:120-124 -> (B,6)."""
        if not self.forces:
            return self.v.new_zeros(self.v.shape[0], 6)
        return sum(f.force(t) for f in self.forces)

    def batch(self):
        return max(t.shape[0] for t in (self.p, self.v, self.mass, self.fric_coeff, self.restitution,
                                        self.ang_inertia, self.scale, self.shape))


_MESH_TENSORS = {}


def _mesh_tensors(verts, faces, device):
    """(verts f64, faces i32) tensors on ``device``.  Read-only numpy meshes (the memoised analytic meshes of meshes.py)
    are uploaded once per device and shared by every body built from them (worlds are rebuilt each optimisation
    iteration; a 20x1x20 floor is 4 MB)."""
    key = None
    if isinstance(verts, np.ndarray) and isinstance(faces, np.ndarray) and not verts.flags.writeable \
            and not faces.flags.writeable:
        key = (id(verts), id(faces), str(device))
        hit = _MESH_TENSORS.get(key)
        if hit is not None and hit[0] is verts:
            return hit[1], hit[2]
    v = verts if isinstance(verts, torch.Tensor) else torch.from_numpy(np.array(verts, dtype=np.float64))
    f = faces if isinstance(faces, torch.Tensor) else torch.from_numpy(np.array(faces))
    v, f = v.to(F64), f.to(torch.int32)
    if device is not None:
        v, f = v.to(device), f.to(device)
    if key is not None:
        if len(_MESH_TENSORS) > 64:
            _MESH_TENSORS.clear()
        _MESH_TENSORS[key] = (verts, v, f)          # keeps the numpy array alive, so its id stays unique
    return v, f


class SDF3D(Body3D):
    """bodies.py:627-760; ``kind`` selects the SDF evaluated on the device."""
    kind = None

    def __init__(self, pos, scale, shape, mesh, sdf_grid=None, **kw):
        device = kw.get('device')
        self.scale = as_batched(scale, 0, device)
        self.shape = as_batched(shape, 1, device)
        # mesh = (verts, faces) or, for per-world topologies, (verts (B,Vmax,3), faces (B,Fmax,3), nverts (B,), nfaces (B,))
        verts, faces = mesh[0], mesh[1]
        self.verts, self.faces = _mesh_tensors(verts, faces, device)
        if len(mesh) == 4:
            self.nverts_w, self.nfaces_w = mesh[2], mesh[3]
        self.sdf_grid = sdf_grid
        super().__init__(pos, **kw)

    def get_surface(self):
        """bodies.py:717-719 -> world-frame vertices (B,V,3), faces."""
        from .transforms import quaternion_apply
        v = self.verts if self.verts.dim() == 3 else self.verts.unsqueeze(0)
        return quaternion_apply(self.rot.unsqueeze(1), v) + self.pos.unsqueeze(1), self.faces

    def shape_rows(self):
        """(B,4) = [a,b,c,scale] rows the kernels read."""
        B = max(self.shape.shape[0], self.scale.shape[0])
        return torch.cat([self.shape.expand(B, 3), self.scale.expand(B).unsqueeze(1)], 1)

    def query_sdfs(self, pts_loc, return_grads=True, return_overlapmask=False):
        """bodies.py:721-760 on the device: pts_loc (N,3) or (B,N,3) in the body frame."""
        single = pts_loc.dim() == 2
        pts = pts_loc.unsqueeze(0) if single else pts_loc
        rows = self.shape_rows().to(pts.device)
        rows = rows.expand(pts.shape[0], 4).contiguous()
        grid = self.sdf_grid.to(pts.device) if self.sdf_grid is not None else None
        if grid is not None and grid.dim() == 4 and grid.shape[0] != pts.shape[0]:
            grid = grid[0]
        out = ops.sdf_query(self.kind, rows, pts, grid, want_dir=return_grads, extra=getattr(self, 'sdf_extra', (0.0, 0.0)))
        sd, gr = out if return_grads else (out, None)
        mask = torch.all(pts.abs() <= rows[:, 3].reshape(-1, 1, 1), dim=2)
        if single:
            sd, gr, mask = sd[0], (gr[0] if gr is not None else None), mask[0]
        res = (sd,) + ((gr,) if return_grads else ()) + ((mask,) if return_overlapmask else ())
        return res if len(res) > 1 else res[0]


def _check_custom(custom_mesh, custom_inertia):
    """The reference's defaults are custom_mesh = custom_inertia = False (physics3d/utils.py:56-57): bodies are meshed by
    128^3 marching cubes (a third-party CUDA extension) and their inertia integrated over that mesh.  Here meshes are
    hot-path INPUTS: the primitives generate the reference's ``custom_mesh=True`` lattices and closed-form inertias unless
    ``mesh=`` is passed.  Asking explicitly for the marching-cubes variants is an error rather than a silent difference."""
    if custom_mesh is False or custom_inertia is False:
        raise NotImplementedError('custom_mesh=False / custom_inertia=False (marching-cubes mesh + mesh-integrated inertia) '
                                  'are not built: pass mesh=(verts, faces) (e.g. meshes.surface_nets of a sampled SDF) and, '
                                  'for grid bodies, inertia=meshes.mesh_inertia(...)')


class SDFBox(SDF3D):
    """bodies.py:778-854."""
    kind = BOX

    def __init__(self, pos, dims, vel=(0, 0, 0, 0, 0, 0), mass=1, restitution=Defaults3D.RESTITUTION,
                 fric_coeff=Defaults3D.FRIC_COEFF, eps=Defaults3D.EPSILON, custom_mesh=True, custom_inertia=True,
                 mesh=None, max_tri_length=0.1, device=None, **_ignored):
        _check_custom(custom_mesh, custom_inertia)
        self.dims = as_batched(dims, 1, device)
        scale = self.dims.max(dim=1)[0] * 1.5 / 2
        if mesh is None:
            if self.dims.shape[0] == 1:
                mesh = meshes.box_mesh(self.dims[0].tolist(), max_tri_length)
                if self.dims.requires_grad:
                    # differentiable lattice (the reference builds its custom mesh from `dims` with torch ops,
                    # bodies.py:799-854): same values (x * (d / d) = x), gradient d vert / d dims = vert / dims
                    v = torch.as_tensor(np.array(mesh[0]), dtype=F64, device=self.dims.device)
                    mesh = (v * (self.dims[0] / self.dims[0].detach()), mesh[1])
            else:   # per-world dims: same lattice topology (from the largest box), vertices scaled per world
                ref = self.dims.max(dim=0)[0]
                v, f = meshes.box_mesh(ref.tolist(), max_tri_length)
                v = torch.as_tensor(v, dtype=F64, device=self.dims.device)
                mesh = (v.unsqueeze(0) * (self.dims / ref).unsqueeze(1), f)
        super().__init__(pos, scale, self.dims / scale.unsqueeze(1), mesh, vel=vel, mass=mass,
                         restitution=restitution, fric_coeff=fric_coeff, eps=eps, device=device)

    def _get_ang_inertia(self, mass):
        d = self.dims
        diag = (d[:, [1, 0, 0]] ** 2 + d[:, [2, 2, 1]] ** 2) / 12
        return mass.reshape(-1, 1, 1) * torch.diag_embed(diag)


class SDFSphere(SDF3D):
    """bodies.py:952-1009 (icosphere, 4 subdivisions)."""
    kind = SPHERE

    def __init__(self, pos, rad, vel=(0, 0, 0, 0, 0, 0), mass=1, restitution=Defaults3D.RESTITUTION,
                 fric_coeff=Defaults3D.FRIC_COEFF, eps=Defaults3D.EPSILON, custom_mesh=True, custom_inertia=True,
                 mesh=None, subdivisions=4, device=None, **_ignored):
        _check_custom(custom_mesh, custom_inertia)
        self.rad = as_batched(rad, 0, device)
        scale = self.rad * 1.5
        if mesh is None:
            v, f = meshes.icosphere(1.0, subdivisions)
            v = torch.as_tensor(v, dtype=F64, device=self.rad.device)
            mesh = ((v * self.rad[0]) if self.rad.shape[0] == 1 else v.unsqueeze(0) * self.rad.reshape(-1, 1, 1), f)
        shape = torch.stack([self.rad / scale, self.rad * 0, self.rad * 0], 1)
        super().__init__(pos, scale, shape, mesh, vel=vel, mass=mass, restitution=restitution,
                         fric_coeff=fric_coeff, eps=eps, device=device)

    def _get_ang_inertia(self, mass):
        return (2 / 5 * mass * self.rad ** 2).reshape(-1, 1, 1) * torch.eye(3, dtype=F64, device=mass.device)


class SDFCylinder(SDF3D):
    """bodies.py:889-949 (axis = local z)."""
    kind = CYLINDER

    def __init__(self, pos, rad, height, vel=(0, 0, 0, 0, 0, 0), mass=1, restitution=Defaults3D.RESTITUTION,
                 fric_coeff=Defaults3D.FRIC_COEFF, eps=Defaults3D.EPSILON, custom_mesh=True, custom_inertia=True,
                 mesh=None, numsegs=32, max_tri_length=0.1, device=None, **_ignored):
        _check_custom(custom_mesh, custom_inertia)
        self.rad, self.height = as_batched(rad, 0, device), as_batched(height, 0, device)
        assert self.rad.shape[0] == 1 and self.height.shape[0] == 1, 'per-world cylinder sizes: pass mesh= explicitly'
        scale = torch.max(self.rad, self.height / 2) * 1.5
        if mesh is None:
            mesh = meshes.cylinder_mesh(float(self.rad[0]), float(self.height[0]), numsegs, max_tri_length)
        shape = torch.stack([self.rad / scale, self.height / scale, self.rad * 0], 1)
        super().__init__(pos, scale, shape, mesh, vel=vel, mass=mass, restitution=restitution,
                         fric_coeff=fric_coeff, eps=eps, device=device)

    def _get_ang_inertia(self, mass):
        r, h = self.rad, self.height
        a = (3 * r ** 2 + h ** 2) / 12
        return mass.reshape(-1, 1, 1) * torch.diag_embed(torch.stack([a, a, r ** 2 / 2], 1))


class SDFGrid3D(SDF3D):
    """bodies.py:763-775: SDF sampled on a res^3 grid over [-1,1]^3 (times scale); mesh handed in."""
    kind = GRID

    def __init__(self, pos, scale, sdf, mesh, vel=(0, 0, 0), mass=1, restitution=Defaults3D.RESTITUTION,
                 fric_coeff=Defaults3D.FRIC_COEFF, eps=Defaults3D.EPSILON, inertia=None, device=None, **_ignored):
        grid = torch.as_tensor(sdf).to(F64)
        if device is not None:
            grid = grid.to(device)
        self._unit_inertia = inertia
        self._mesh_np = mesh
        sc = as_batched(scale, 0, device)
        super().__init__(pos, sc, torch.zeros(1, 3, dtype=F64), mesh, sdf_grid=grid, vel=vel, mass=mass,
                         restitution=restitution, fric_coeff=fric_coeff, eps=eps, device=device)

    def _get_ang_inertia(self, mass):
        if self._unit_inertia is not None:
            I = torch.as_tensor(self._unit_inertia, dtype=F64, device=mass.device)
        else:
            v = self.verts[0] if self.verts.dim() == 3 else self.verts
            I = torch.as_tensor(meshes.mesh_inertia(v.cpu().numpy(), self.faces.cpu().numpy()), dtype=F64,
                                device=mass.device)
        return mass.reshape(-1, 1, 1) * I


class SDFDecoder3D(SDFGrid3D):
    """A body whose shape is a neural SDF ``sdf_func(points (N,3), *params) -> (N,)`` on [-1,1]^3 (times ``scale``) -- the
    reference's IGR-decoder bodies (SDF3D with ``sdf_func=decode_igr(network)``, ``params=[latent]``, bodies.py:627-760,
    physics3d/utils.py:330-350) in the form BASELINE's config 4 prescribes for this path: the decoder is BAKED to a res^3
    grid that the contact kernels query (``igr.bake_grid`` semantics), while mesh and inertia stay differentiable
    functions of ``params`` exactly as in the reference (``MeshSDF`` backward, bodies.py:652-704; volume-integral
    inertia, :260-395).  Gradients therefore reach ``params`` through the mesh vertices and the inertia; the baked SDF
    values themselves carry none (the reference's own ``DiffGridSDF`` gives grid values no gradient either, :246-257)."""

    def __init__(self, pos, scale, sdf_func, params, res=64, mesh_res=None, device=None, **kw):
        params = [q if isinstance(q, torch.Tensor) else torch.as_tensor(q, dtype=F64) for q in params]
        t = torch.linspace(-1.0, 1.0, res, dtype=F64)
        pts = torch.stack(torch.meshgrid(t, t, t, indexing='ij'), 3).reshape(-1, 3)
        with torch.no_grad():
            grid = torch.cat([sdf_func(pts[i:i + (1 << 16)], *[q.detach().cpu() for q in params])
                              for i in range(0, pts.shape[0], 1 << 16)]).reshape(res, res, res)
        verts, faces = meshes.iso_surface_mesh(sdf_func, params, res=mesh_res or res)
        sc = float(scale)
        inertia = meshes.mesh_inertia_torch(verts * sc, faces)
        if device is not None:
            verts, faces, inertia = verts.to(device), faces.to(device), inertia.to(device)
        self.sdf_params = params
        self.sdf_func = sdf_func
        super().__init__(pos, scale, grid, mesh=(verts * sc, faces), inertia=inertia, device=device, **kw)

    def query_sdfs(self, pts_loc, return_grads=True, return_overlapmask=False, exact=False):
        """``exact=False`` (default): the baked grid through the CUDA query operator, like every kernel of the stepping
        path.  ``exact=True``: the decoder itself, evaluated in torch exactly as the reference's ``SDF3D.query_sdfs`` does
        for bodies without a closed-form gradient (bodies.py:721-760): value = ``sdf_func(p / scale) * scale`` inside the
        cube (``scale`` outside), direction = normalised autograd gradient; the values stay attached to ``params``
        (detached when directions are requested for points that do not require grad, :744-745) -- what the point-cloud
        fitting loss differentiates (``losses.pointcloud_sdf_loss(..., exact=True)``).  Single world: pts_loc (N,3)."""
        if not exact:
            return super().query_sdfs(pts_loc, return_grads=return_grads, return_overlapmask=return_overlapmask)
        scale = self.scale.reshape(-1)[0].to(pts_loc.device)
        params = [q.to(pts_loc.device) for q in self.sdf_params]
        mask = torch.all(pts_loc.abs() <= scale, dim=1)
        sdfs = torch.ones(pts_loc.shape[0], dtype=pts_loc.dtype, device=pts_loc.device)
        grads = torch.zeros_like(pts_loc)
        if bool(mask.any()):
            pts_in = pts_loc[mask] / scale
            if return_grads:
                with torch.enable_grad():
                    leaf = pts_in.is_leaf
                    if not pts_in.requires_grad:
                        pts_in.requires_grad_(True)
                    vals = self.sdf_func(pts_in, *params)
                    g = torch.autograd.grad(vals, pts_in, torch.ones_like(vals), retain_graph=not leaf)[0]
                sdfs = sdfs.clone()
                sdfs[mask] = vals
                grads[mask] = torch.nn.functional.normalize(g, dim=1)
                if leaf:
                    sdfs = sdfs.detach()
            else:
                sdfs = sdfs.clone()
                sdfs[mask] = self.sdf_func(pts_in, *params)
        sdfs = sdfs * scale
        res = (sdfs,) + ((grads,) if return_grads else ()) + ((mask,) if return_overlapmask else ())
        return res if len(res) > 1 else res[0]


def _sampled_mesh(kind, shape3, extra, scale, res=96):
    """Iso-surface mesh of an analytic SDF body sampled on a res^3 lattice over its cube (what the reference extracts with
    128^3 marching cubes when custom_mesh=False, bodies.py:652-712): meshes.surface_nets of meshes.sample_sdf."""
    g = meshes.sample_sdf(kind, [float(x) for x in shape3], [float(x) for x in extra], res)
    v, f = meshes.surface_nets(g)
    return v * float(scale), f


class SDFBoxRounded(SDF3D):
    """bodies.py:856-870: box_sdf(dims - 2 r) - r.  The mesh is the iso-surface of the sampled SDF unless given."""
    kind = BOX_ROUNDED

    def __init__(self, pos, dims, r, vel=(0, 0, 0, 0, 0, 0), mass=1, restitution=Defaults3D.RESTITUTION,
                 fric_coeff=Defaults3D.FRIC_COEFF, eps=Defaults3D.EPSILON, mesh=None, inertia=None, device=None, **_ignored):
        self.dims = as_batched(dims, 1, device)
        assert self.dims.shape[0] == 1, 'per-world dims are not supported for rounded boxes'
        self.r = float(r)
        scale = self.dims.max(dim=1)[0] * 1.5 / 2
        shape = (self.dims - 2 * self.r) / scale.unsqueeze(1)
        self.sdf_extra = (self.r / float(scale[0]), 0.0)
        if mesh is None:
            mesh = _sampled_mesh('box_rounded', shape[0].tolist(), self.sdf_extra, scale[0])
        self._unit_inertia = inertia
        super().__init__(pos, scale, shape, mesh, vel=vel, mass=mass, restitution=restitution, fric_coeff=fric_coeff,
                         eps=eps, device=device)

    def _get_ang_inertia(self, mass):
        I = self._unit_inertia
        if I is None:
            I = meshes.mesh_inertia(self.verts.detach().cpu().numpy(), self.faces.cpu().numpy())
        return mass.reshape(-1, 1, 1) * torch.as_tensor(I, dtype=F64, device=mass.device)


class SDFBrick(SDFBoxRounded):
    """bodies.py:873-886: box whose first two dimensions are rounded in-plane (brick_sdf); the reference pairs it with the
    direction function of a cube of side r (its rounded_sdf_grad wrapper drops the first parameter), reproduced here."""
    kind = BRICK

    def __init__(self, pos, dims, r, vel=(0, 0, 0, 0, 0, 0), mass=1, restitution=Defaults3D.RESTITUTION,
                 fric_coeff=Defaults3D.FRIC_COEFF, eps=Defaults3D.EPSILON, mesh=None, inertia=None, device=None, **_ignored):
        self.dims = as_batched(dims, 1, device)
        assert self.dims.shape[0] == 1, 'per-world dims are not supported for bricks'
        self.r = float(r)
        scale = self.dims.max(dim=1)[0] * 1.5 / 2
        shape = self.dims / scale.unsqueeze(1)
        self.sdf_extra = (self.r / float(scale[0]), 0.0)
        if mesh is None:
            mesh = _sampled_mesh('brick', shape[0].tolist(), self.sdf_extra, scale[0])
        self._unit_inertia = inertia
        SDF3D.__init__(self, pos, scale, shape, mesh, vel=vel, mass=mass, restitution=restitution, fric_coeff=fric_coeff,
                       eps=eps, device=device)


class SDFBowl(SDFBoxRounded):
    """bodies.py:1012-1026: half shell of radius r and half thickness d (bowl_sdf / bowl_sdf_grad), scale (r + d) * 1.3333."""
    kind = BOWL

    def __init__(self, pos, r, d, vel=(0, 0, 0), mass=1, restitution=Defaults3D.RESTITUTION,
                 fric_coeff=Defaults3D.FRIC_COEFF, eps=Defaults3D.EPSILON, mesh=None, inertia=None, device=None, **_ignored):
        self.r, self.d = float(r), float(d)
        scale = as_batched((self.r + self.d) * 1.3333, 0, device)
        shape = torch.tensor([[self.r, self.d, 0.0]], dtype=F64, device=scale.device) / scale.unsqueeze(1)
        self.sdf_extra = (0.0, 0.0)
        if mesh is None:
            mesh = _sampled_mesh('bowl', shape[0].tolist(), self.sdf_extra, scale[0])
        self._unit_inertia = inertia
        SDF3D.__init__(self, pos, scale, shape, mesh, vel=vel, mass=mass, restitution=restitution, fric_coeff=fric_coeff,
                       eps=eps, device=device)
