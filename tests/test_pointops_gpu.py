"""SDF query and integrator kernels vs reference golden vectors and the oracle's autograd."""
import os

import numpy as np
import pytest
import torch

from diffsdfsim_b200 import scenes
from oracle import sdf as S, transforms as T
from oracle.scenes import shape_for

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')
F64 = torch.float64


def _body(kind):
    spec = {'box': scenes.box_on_plane()['bodies'][1], 'sphere': scenes.bouncing_sphere()['bodies'][1],
            'cylinder': scenes.grid_on_pole()['bodies'][1], 'grid': scenes.grid_on_pole()['bodies'][2]}[kind]
    k, params, scale, _ = shape_for(spec, torch.tensor(1.0, dtype=F64))
    return k, params, scale


def _shape_row(kind, params, scale):
    row = torch.zeros(4, dtype=F64)
    if kind != 'grid':
        flat = torch.cat([p.reshape(-1) for p in params])
        row[:flat.numel()] = flat
    row[3] = scale
    return row


@pytest.mark.parametrize('kind', ['box', 'sphere', 'cylinder', 'grid'])
def test_query_matches_reference_golden(kind):
    from diffsdfsim_b200.ops import sdf_query
    g = np.load(os.path.join(GOLD, 'sdf_query.npz'))
    k, params, scale = _body(kind)
    pts = torch.tensor(g[kind + '_pts'], device='cuda')[None]
    shape = _shape_row(kind, params, scale).cuda()[None]
    grid = params[0].cuda() if kind == 'grid' else None
    sd, d = sdf_query(kind, shape, pts, grid)
    # same arithmetic, different instruction selection (fma contraction): a few ulp
    np.testing.assert_allclose(sd[0].cpu().numpy(), g[kind + '_sdf'], rtol=0, atol=1e-14)
    np.testing.assert_allclose(d[0].cpu().numpy(), g[kind + '_dir'], rtol=0, atol=1e-13)
    # in/out-of-cube classification and zero-direction set are bit-exact
    assert np.array_equal(np.all(d[0].cpu().numpy() == 0, axis=1), np.all(g[kind + '_dir'] == 0, axis=1))


@pytest.mark.parametrize('kind', ['box', 'sphere', 'cylinder', 'grid'])
def test_query_backward_matches_oracle_autograd(kind):
    from diffsdfsim_b200.ops import sdf_query
    k, params, scale = _body(kind)
    gen = torch.Generator().manual_seed(1)
    W, N = 2, 2000
    pts = (torch.rand(W, N, 3, generator=gen, dtype=F64) * 2 - 1) * float(scale) * 1.05
    pts[:, :32] = torch.round(pts[:, :32] / float(scale) * 4) / 4 * float(scale)      # exact zeros / ties
    ws, wd = torch.randn(W, N, generator=gen, dtype=F64), torch.randn(W, N, 3, generator=gen, dtype=F64)
    po = pts.clone().requires_grad_(True)
    tot = 0.
    for w in range(W):
        sd, d = S.query(k, params, scale, po[w])
        tot = tot + (sd * ws[w]).sum() + (d * wd[w]).sum()
    tot.backward()
    pg = pts.cuda().requires_grad_(True)
    shape = _shape_row(kind, params, scale).cuda()[None].repeat(W, 1)
    grid = params[0].cuda()[None].repeat(W, 1, 1, 1) if kind == 'grid' else None
    sd, d = sdf_query(kind, shape, pg, grid)
    ((sd * ws.cuda()).sum() + (d * wd.cuda()).sum()).backward()
    ref = po.grad.numpy()
    np.testing.assert_allclose(pg.grad.cpu().numpy(), ref, rtol=1e-9, atol=1e-9 * max(1.0, np.abs(ref).max()))


def test_integrator_forward_backward_vs_oracle():
    from diffsdfsim_b200.ops import integrate
    gen = torch.Generator().manual_seed(2)
    W, nb = 64, 3
    q = torch.randn(W, nb, 4, generator=gen, dtype=F64)
    q = q / q.norm(dim=-1, keepdim=True)
    q[0, 0] = torch.tensor([1., 0, 0, 0])
    q[1, 0] = torch.tensor([0., 1, 0, 0])            # forces a non-zero matrix_to_quaternion branch
    p = torch.cat([q, torch.randn(W, nb, 3, generator=gen, dtype=F64)], -1)
    v = torch.randn(W, nb, 6, generator=gen, dtype=F64) * 2
    v[2] *= 1e-3                                      # |w dt| below the 1e-2 clamp of so3_exponential_map
    v[3, :, :3] = 0
    v[4, :, :3] *= 40                                 # large rotation: other quaternion candidates
    dt = torch.rand(W, generator=gen, dtype=F64) * 0.05 + 1e-3
    wgt = torch.randn(W, nb, 7, generator=gen, dtype=F64)

    po, vo, dto = p.clone().requires_grad_(True), v.clone().requires_grad_(True), dt.clone().requires_grad_(True)
    out = []
    for w in range(W):
        for b in range(nb):
            dq = T.matrix_to_quaternion(T.so3_exponential_map(vo[w, b, :3].unsqueeze(0) * dto[w]))
            out.append(torch.cat([T.quaternion_multiply(dq, po[w, b, :4]).squeeze(), po[w, b, 4:] + vo[w, b, 3:] * dto[w]]))
    out = torch.stack(out).reshape(W, nb, 7)
    (out * wgt).sum().backward()

    pg, vg, dtg = [t.cuda().requires_grad_(True) for t in (p, v, dt)]
    res = integrate(pg, vg, dtg)
    (res * wgt.cuda()).sum().backward()
    np.testing.assert_allclose(res.detach().cpu().numpy(), out.detach().numpy(), rtol=0, atol=1e-14)
    for a, b in ((pg, po), (vg, vo), (dtg, dto)):
        np.testing.assert_allclose(a.grad.cpu().numpy(), b.grad.numpy(), rtol=1e-10, atol=1e-11)
    # masked worlds pass through
    act = torch.ones(W, dtype=torch.uint8, device='cuda')
    act[::2] = 0
    res2 = integrate(pg.detach(), vg.detach(), dtg.detach(), act)
    assert torch.equal(res2[::2], pg.detach()[::2]) and torch.equal(res2[1::2], res.detach()[1::2])


@pytest.mark.parametrize('kind', ['box_rounded', 'brick', 'bowl'])
def test_extra_sdf_kinds_match_reference_golden_and_oracle_autograd(kind):
    """Rounded box / brick / bowl (bodies.py:128-200): values and directions vs the reference instances, point gradients
    vs the oracle's autograd."""
    from diffsdfsim_b200.ops import sdf_query
    from specs import extra_sdf_kinds
    g = np.load(os.path.join(GOLD, 'sdf_query.npz'))
    k, params, scale, row, extra = extra_sdf_kinds()[kind]
    pts = torch.tensor(g[kind + '_pts'], device='cuda')[None]
    sd, d = sdf_query(kind, row.cuda()[None], pts, extra=extra)
    np.testing.assert_allclose(sd[0].cpu().numpy(), g[kind + '_sdf'], rtol=0, atol=1e-14)
    np.testing.assert_allclose(d[0].cpu().numpy(), g[kind + '_dir'], rtol=0, atol=1e-13)
    gen = torch.Generator().manual_seed(1)
    N = 2000
    p0 = (torch.rand(N, 3, generator=gen, dtype=F64) * 2 - 1) * float(scale) * 1.05
    ws, wd = torch.randn(N, generator=gen, dtype=F64), torch.randn(N, 3, generator=gen, dtype=F64)
    po = p0.clone().requires_grad_(True)
    so, do = S.query(k, params, scale, po)
    ((so * ws).sum() + (do * wd).sum()).backward()
    pg = p0.cuda()[None].requires_grad_(True)
    sg, dg = sdf_query(kind, row.cuda()[None], pg, extra=extra)
    ((sg[0] * ws.cuda()).sum() + (dg[0] * wd.cuda()).sum()).backward()
    ref = po.grad.numpy()
    np.testing.assert_allclose(pg.grad[0].cpu().numpy(), ref, rtol=1e-9, atol=1e-9 * max(1.0, np.abs(ref).max()))
