"""ORACLE (test infrastructure, not product code) -- signed-distance evaluation.

CPU float64 torch restatement of the reference's SDF operators, written so that
autograd through these functions yields the same derivative conventions as the
reference (clamp vs. maximum sub-gradients, sign(0) handling):

* box / sphere / cylinder value + "failsafe" gradient
  -> sdf_physics/physics3d/bodies.py:38-125
* grid SDF (trilinear) + central-difference gradient field + custom backward
  -> sdf_physics/physics3d/bodies.py:203-257
* ``query`` (inside-cube mask, p/scale, value*scale, unit gradient)
  -> sdf_physics/physics3d/bodies.py:721-760
* ``grid_interp``: third-party ``ev_sdf_utils.grid_interp`` (un-vendored,
  unpinned; call sites bodies.py:209,241) restated from SURVEY.md Appendix A:
  trilinear at fractional index coordinates, base cell clamped to R-2.
"""
import torch
from torch.nn.functional import normalize

BOX, SPHERE, CYLINDER, GRID, BOX_ROUNDED, BRICK, BOWL = 0, 1, 2, 3, 4, 5, 6


# ----------------------------------------------------------------- analytic
def box_value(x, dims):
    q = x.abs() - dims / 2
    return q.clamp(min=0.).norm(dim=1) + q.max(dim=1)[0].clamp(max=0.)


def box_direction(x, dims):
    q = x.abs() - dims / 2
    sgn = torch.sign(x)
    sgn[sgn == 0] = 1
    top = q.max(dim=1)[0]
    tie = torch.zeros_like(x)
    tie[q == top.unsqueeze(1)] = 1.            # diagonal failsafe on edges/corners inside
    outside = torch.max(q, q.new_zeros(1))
    g = (normalize(outside, dim=1) + (top <= 0).to(x.dtype).unsqueeze(1) * tie) * sgn
    return normalize(g, dim=1)


def sphere_value(x, rad):
    return x.norm(dim=1) - rad


def sphere_direction(x, rad):
    return normalize(x, dim=1)


def _cyl_q(x, rad, height):
    rz = torch.stack([x[:, :2].norm(dim=1), x[:, 2]], dim=1)
    return rz.abs() - torch.stack([rad, height / 2])


def cylinder_value(x, rad, height):
    q = _cyl_q(x, rad, height)
    return q.clamp(min=0.).norm(dim=1) + q.max(dim=1)[0].clamp(max=0.)


def cylinder_direction(x, rad, height):
    q = _cyl_q(x, rad, height)
    sgn = torch.sign(x[:, 2])
    sgn[sgn == 0] = 1
    top = q.max(dim=1)[0]
    tie = torch.zeros_like(q)
    tie[q == top.unsqueeze(1)] = 1.
    g2 = normalize(q.clamp(min=0.), dim=1) + (top <= 0).to(x.dtype).unsqueeze(1) * tie
    g = torch.cat([g2[:, 0:1] * normalize(x[:, :2], dim=1), (g2[:, 1] * sgn).unsqueeze(1)], dim=1)
    return normalize(g, dim=1)


# ------------------------------------------------- rounded box / brick / bowl (bodies.py:128-200)
def rounded_box_value(x, dims, r):
    """rounded_sdf(box_sdf) with params [r, dims] (bodies.py:166-172, 866-867)."""
    return box_value(x, dims) - r


def rounded_box_direction(x, dims, r):
    return box_direction(x, dims)


def brick_value(x, dims, r):
    """bodies.py:184-200."""
    half = torch.cat([dims[:2] / 2 - r, dims[2:3] / 2])
    q = x.abs() - half
    top01 = q[:, :2].max(dim=1)[0]
    s01 = q[:, :2].clamp(min=0.).norm(dim=1) + top01.clamp(max=0.) - r
    q2 = torch.stack([s01, q[:, 2]], dim=1)
    return q2.clamp(min=0.).norm(dim=1) + q2.max(dim=1)[0].clamp(max=0.)


def brick_direction(x, dims, r):
    """The reference's wrapper hands the SECOND parameter (r) to box_sdf_grad as the box size (bodies.py:175-181, 882-884)."""
    return box_direction(x, r)


def _bowl_ps(x, r, d):
    ps = torch.stack([x[:, :2].norm(dim=1), x[:, 2]], dim=1)
    nrm = ps.norm(dim=1)
    first = torch.where(ps[:, 1] < 0, nrm, ps[:, 0])
    return torch.stack([(first - r).abs() - d, ps[:, 1]], dim=1), nrm


def bowl_value(x, r, d):
    """bodies.py:128-142; the in-place shift of the caller's tensor is made explicit by ``query``."""
    ps, _ = _bowl_ps(x, r, d)
    return torch.max(ps, ps.new_zeros(1)).norm(dim=1) + torch.min(ps.new_zeros(1), ps.max(dim=1)[0])


def bowl_direction(x, r, d):
    """bodies.py:145-163."""
    ps, nrm = _bowl_ps(x, r, d)
    g = x * (nrm - r).sign().unsqueeze(1)
    up = ps[:, 1] >= 0
    gxy = torch.where((up & (ps[:, 0] < 0)).unsqueeze(1), torch.zeros_like(g[:, :2]), g[:, :2])
    gz = torch.where(up, g[:, 2].abs(), g[:, 2])
    return normalize(torch.cat([gxy, gz.unsqueeze(1)], dim=1), dim=1)


# --------------------------------------------------------------------- grid
def grid_interp(grid, idx):
    """Trilinear interpolation at fractional INDEX coordinates.

    grid (R0,R1,R2) -> (N,);  channel-first grid (C,R0,R1,R2) -> (N,C).
    Callers guarantee 0 <= idx <= R-1; base cell clamped to R-2 so idx == R-1
    evaluates on the last cell with weight 1.  No gradient w.r.t. idx.
    """
    vec = grid.dim() == 4
    g = grid if vec else grid.unsqueeze(0)
    dims = torch.tensor(g.shape[1:], dtype=torch.long)
    idx = idx.detach()
    base = torch.minimum(torch.floor(idx).long().clamp(min=0), (dims - 2).unsqueeze(0))
    t = idx - base.to(idx.dtype)
    out = g.new_zeros((g.shape[0], idx.shape[0]))
    for dx in (0, 1):
        wx = t[:, 0] if dx else 1 - t[:, 0]
        for dy in (0, 1):
            wy = t[:, 1] if dy else 1 - t[:, 1]
            for dz in (0, 1):
                wz = t[:, 2] if dz else 1 - t[:, 2]
                out = out + g[:, base[:, 0] + dx, base[:, 1] + dy, base[:, 2] + dz] * (wx * wy * wz)
    return out.t() if vec else out[0]


def _grid_index(x, grid):
    ext = x.new_tensor(grid.shape) - 1
    idx = (x + 1.) * 0.5 * ext
    ok = torch.all((idx <= ext) & (idx >= 0), dim=1)
    return idx, ok


def grid_value_raw(x, grid):
    idx, ok = _grid_index(x, grid)
    out = grid.new_ones(x.shape[0])
    out[ok] = grid_interp(grid, idx[ok])
    return out


def central_difference_field(grid):
    """(3,R,R,R) central differences in index units, zero on both boundary planes of each axis."""
    f = grid.new_zeros((3,) + tuple(grid.shape))
    f[0, 1:-1] = (grid[2:] - grid[:-2]) / 2
    f[1, :, 1:-1] = (grid[:, 2:] - grid[:, :-2]) / 2
    f[2, :, :, 1:-1] = (grid[:, :, 2:] - grid[:, :, :-2]) / 2
    return f


def grid_direction(x, grid):
    idx, ok = _grid_index(x, grid)
    out = grid.new_zeros(x.shape)
    out[ok] = normalize(grid_interp(central_difference_field(grid), idx[ok]), dim=1)
    return out


class _GridValue(torch.autograd.Function):
    """Value = trilinear; d value/d x := unit interpolated central-difference direction."""

    @staticmethod
    def forward(ctx, x, grid):
        ctx.save_for_backward(x, grid)
        ctx.mark_non_differentiable(grid)
        return grid_value_raw(x, grid)

    @staticmethod
    def backward(ctx, g):
        x, grid = ctx.saved_tensors
        return grid_direction(x, grid) * g.unsqueeze(1), None


def grid_value(x, grid):
    return _GridValue.apply(x, grid)


_VALUE = {BOX: box_value, SPHERE: sphere_value, CYLINDER: cylinder_value, GRID: grid_value,
          BOX_ROUNDED: rounded_box_value, BRICK: brick_value, BOWL: bowl_value}
_DIRECTION = {BOX: box_direction, SPHERE: sphere_direction, CYLINDER: cylinder_direction, GRID: grid_direction,
              BOX_ROUNDED: rounded_box_direction, BRICK: brick_direction, BOWL: bowl_direction}


def query(kind, params, scale, pts, want_dir=True):
    """SDF3D.query_sdfs: pts in the body frame -> (sdf, unit gradient).

    Outside the cube |p| <= scale: sdf = 1*scale, gradient 0.
    ``params`` are the normalised shape parameters (already divided by scale).
    """
    inside = torch.all(pts.abs() <= scale, dim=1)
    val = pts.new_ones(pts.shape[0])
    direction = pts.new_zeros(pts.shape)
    if torch.any(inside):
        u = pts[inside] / scale
        if kind == BOWL:
            # bowl_sdf and bowl_sdf_grad both shift the z coordinate of the SAME tensor in place (bodies.py:129,146):
            # the value sees z - r/2, the direction z - r
            shift = torch.stack([torch.zeros_like(params[0]), torch.zeros_like(params[0]), params[0] / 2])
            u = u - shift
            val[inside] = _VALUE[kind](u, *params)
            if want_dir:
                direction[inside] = normalize(_DIRECTION[kind](u - shift, *params), dim=1)
        else:
            val[inside] = _VALUE[kind](u, *params)
            if want_dir:
                direction[inside] = normalize(_DIRECTION[kind](u, *params), dim=1)
    val = val * scale
    return (val, direction) if want_dir else val
