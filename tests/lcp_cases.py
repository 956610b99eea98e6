"""Contact-structured random LCP instances (engines.py:56-79 shape) for solver parity tests."""
import numpy as np
import torch

F64 = torch.float64


def tangent_dirs(n):
    k = int(np.argmin(np.abs(n)))
    e = np.zeros(3)
    e[k] = 1
    d1 = np.cross(e, n); d1 /= np.linalg.norm(d1)
    d2 = np.cross(d1, n); d2 /= np.linalg.norm(d2)
    d3 = d1 + d2; d3 /= np.linalg.norm(d3)
    d4 = np.cross(d3, n); d4 /= np.linalg.norm(d4)
    D = np.stack([d1, d2, d3, d4])
    return np.concatenate([D, -D])


def contact_lcp(rng, nc, nb=2, pinned=(0,), cap=None):
    """Returns dict of numpy arrays Q,p,G,h,A,b,F with nineq = 10*nc (padded with zero rows to 10*cap)."""
    nz = 6 * nb
    M = np.zeros((nz, nz))
    for i in range(nb):
        a = rng.rand(3, 3)
        I = a @ a.T + 0.3 * np.eye(3)
        m = 0.5 + rng.rand()
        M[6 * i:6 * i + 3, 6 * i:6 * i + 3] = I * m * 0.2
        M[6 * i + 3:6 * i + 6, 6 * i + 3:6 * i + 6] = np.eye(3) * m
    v = rng.randn(nz) * 0.5
    for i in pinned:
        v[6 * i:6 * i + 6] = 0
    f = rng.randn(nz)
    f[4::6] -= 10.0
    u = M @ v + f / 30
    Jc = np.zeros((nc, nz)); Jf = np.zeros((8 * nc, nz))
    for c in range(nc):
        i1, i2 = 0, 1 + rng.randint(nb - 1)
        n = np.array([0., 1, 0]) + 0.2 * rng.randn(3); n /= np.linalg.norm(n)
        p1 = rng.randn(3); p2 = rng.randn(3) * 0.5; p2[1] = -0.5
        Jc[c, 6 * i1:6 * i1 + 6] = np.concatenate([np.cross(p1, n), n])
        Jc[c, 6 * i2:6 * i2 + 6] = -np.concatenate([np.cross(p2, n), n])
        D = tangent_dirs(n)
        Jf[8 * c:8 * c + 8, 6 * i1:6 * i1 + 6] = np.concatenate([np.cross(np.tile(p1, (8, 1)), D), D], 1)
        Jf[8 * c:8 * c + 8, 6 * i2:6 * i2 + 6] = -np.concatenate([np.cross(np.tile(p2, (8, 1)), D), D], 1)
    ni = 10 * nc
    G = np.concatenate([Jc, Jf, np.zeros((nc, nz))])
    Fm = np.zeros((ni, ni))
    mu = 0.05 + 0.5 * rng.rand(nc)
    for c in range(nc):
        Fm[nc + 8 * c:nc + 8 * c + 8, 9 * nc + c] = 1
        Fm[9 * nc + c, c] = mu[c]
        Fm[9 * nc + c, nc + 8 * c:nc + 8 * c + 8] = -1
    h = np.concatenate([(Jc @ v) * 0.5, np.zeros(9 * nc)])
    neq = 6 * len(pinned)
    A = np.zeros((neq, nz))
    for r, i in enumerate(pinned):
        A[6 * r:6 * r + 6, 6 * i:6 * i + 6] = np.eye(6)
    out = dict(Q=M, p=u, G=G, h=h, A=A, b=np.zeros(neq), F=Fm)
    if cap is not None and cap > nc:
        nic = 10 * cap
        Gp = np.zeros((nic, nz)); Gp[:ni] = G
        hp = np.zeros(nic); hp[:ni] = h
        Fp = np.zeros((nic, nic)); Fp[:ni, :ni] = Fm
        out.update(G=Gp, h=hp, F=Fp)
    out['nineq'] = ni
    return out


def to_torch(d, device='cpu'):
    return {k: (torch.tensor(v, dtype=F64, device=device) if isinstance(v, np.ndarray) else v) for k, v in d.items()}
