"""``_filter_contacts`` (sdf_physics/physics3d/contacts.py:97-158) as a stand-alone device operator vs the oracle's
restatement of it (scipy's Qhull, like the reference): identical kept SETS on planar, linear, duplicate-ridden and
genuinely three-dimensional contact clusters, including the degenerate inputs Qhull sees in practice (lattice points on
the faces and edges of a box: coplanar facets, collinear edge points)."""
import numpy as np
import pytest
import torch

from diffsdfsim_b200 import ops
from oracle.sim import hull_filter

pytestmark = pytest.mark.gpu
F64 = torch.float64


def _rot(rng, mag):
    w = rng.normal(size=3) * mag
    th = np.linalg.norm(w)
    if th == 0:
        return np.eye(3)
    k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def _cases():
    rng = np.random.default_rng(0)
    up = np.array([0.0, 1.0, 0.0])
    cases = []
    # lattice on the surface of a box: 152 points, hull = 8 corners; every face coplanar, every edge collinear
    g = np.linspace(-0.5, 0.5, 6)
    X, Y, Z = np.meshgrid(g, g * 0.4, g * 0.7, indexing='ij')
    box = np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1)
    surf = box[(np.abs(box[:, 0]) == 0.5) | (np.abs(box[:, 1]) == 0.2) | (np.abs(box[:, 2]) == 0.35)]
    for mag in (0.0, 1e-6, 1e-3, 0.3, 1.0, 2.0):
        cases.append(('box lattice rot %g' % mag, surf @ _rot(rng, mag).T + rng.normal(size=3) * 0.2, None))
    # a resting rounded box: flat bottom patch + a chamfer ring slightly above it (the shape of the scene states)
    u = np.linspace(-0.27, 0.27, 13)
    bx, bz = np.meshgrid(u, u * 0.66, indexing='ij')
    bottom = np.stack([bx.ravel(), np.full(bx.size, -0.25), bz.ravel()], 1)
    ring = np.array([[sx * 0.31, -0.2494, z] for sx in (-1, 1) for z in np.linspace(-0.178, 0.178, 9)])
    for mag in (0.0, 1e-4, 1e-2):
        cases.append(('rounded box patch rot %g' % mag, np.concatenate([bottom, ring]) @ _rot(rng, mag).T, None))
    # random clouds (general position) and points on a sphere (all of them vertices)
    for n in (5, 9, 40, 300):
        cases.append(('cloud %d' % n, rng.normal(size=(n, 3)), None))
    s = rng.normal(size=(120, 3))
    cases.append(('sphere', s / np.linalg.norm(s, axis=1)[:, None], None))
    # planar (2-D path), with interior points and exact duplicates; collinear (1-D path); a single point twice
    pl = np.stack([rng.uniform(-1, 1, 80), np.zeros(80), rng.uniform(-1, 1, 80)], 1)
    cases.append(('planar', np.concatenate([pl, pl[:7]]), None))
    cases.append(('planar tilted', pl @ _rot(rng, 0.7).T + 0.3, None))
    t = rng.uniform(-1, 1, 30)
    cases.append(('collinear', np.stack([t, 0 * t, 2 * t], 1), None))
    cases.append(('short line', np.array([[0.0, 0, 0], [1e-4, 0, 0], [5e-4, 0, 0]]), None))
    # three normal clusters interleaved + zero normals (dropped)
    pts = rng.normal(size=(90, 3))
    nrm = np.zeros((90, 3))
    nrm[0::3] = [0, 1, 0]; nrm[1::3] = [1, 0, 0]; nrm[2::3] = _rot(rng, 0.004) @ up     # third cluster: 0.004 rad from the first
    nrm[5] = 0.0; nrm[17] = 0.0
    cases.append(('three clusters', pts, nrm))
    return [(name, p, (np.tile(up, (len(p), 1)) if n is None else n)) for name, p, n in cases]


def test_filter_matches_oracle_qhull_sets():
    cases = _cases()
    K = max(len(p) for _, p, _ in cases)
    W = len(cases)
    P = torch.zeros(W, K, 3, dtype=F64); N = torch.zeros(W, K, 3, dtype=F64)
    cnt = torch.zeros(W, dtype=torch.int32)
    for w, (_, p, n) in enumerate(cases):
        P[w, :len(p)] = torch.as_tensor(p); N[w, :len(p)] = torch.as_tensor(n); cnt[w] = len(p)
    mask, status = ops.filter_contacts(N.cuda(), P.cuda(), eps=1e-3, count=cnt.cuda())
    mask, status = mask.cpu(), status.cpu()
    assert not bool((status & 4).any()), 'a cluster was too large for the device 3-D hull'
    for w, (name, p, n) in enumerate(cases):
        want = hull_filter(torch.as_tensor(n), torch.as_tensor(p), 1e-3)
        ref_pts = {tuple(p[i]) for i in want.tolist()}
        got = mask[w, :len(p)].nonzero().flatten().tolist()
        got_pts = {tuple(p[i]) for i in got}
        # Qhull returns ONE of several exact duplicates; which index is not specified: compare the kept POINTS
        assert got_pts == ref_pts, f'{name}: kept points differ ({len(got_pts)} vs {len(ref_pts)})'
        assert len(got) == len(want), f'{name}: kept count {len(got)} vs {len(want)} (duplicates must be kept once)'
