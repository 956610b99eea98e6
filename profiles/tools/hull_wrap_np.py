"""numpy model of the device gift wrapping (dsdf_contacts.cu hull3d_vertices) to tune against Qhull."""
import pickle, sys, numpy as np
from scipy.spatial import ConvexHull
TOL1 = float(sys.argv[1]) if len(sys.argv) > 1 else 2e-15
TOL2 = float(sys.argv[2]) if len(sys.argv) > 2 else 2e-15
CLOSE = int(sys.argv[3]) if len(sys.argv) > 3 else 8

def unit(v): return v / np.linalg.norm(v)

def wrap(P, S, T, n0, m0, skip):
    E = T - S; L = np.linalg.norm(E); e = E / L
    d = P - S; al = d @ e; r = d - np.outer(al, e); rho = np.linalg.norm(r, axis=1)
    ok = rho > 1e-14 * (L + np.abs(al))
    for s in skip:
        if s >= 0: ok[s] = False
    if not ok.any(): return -1
    x = -(r @ m0); y = np.maximum(-(r @ n0), 0.0)
    idx = np.nonzero(ok)[0]
    # pass 1: smallest phi -- strict
    best = idx[0]
    for i in idx[1:]:
        det = x[i] * y[best] - x[best] * y[i]
        if det > 0 or (det == 0 and x[i] * x[best] + y[i] * y[best] < 0 and x[i] > 0): best = i
    for it in range(CLOSE):
        det = x[best] * y[idx] - x[idx] * y[best]
        tie = idx[(np.abs(det) <= TOL1 * MAXABS * (rho[best] + rho[idx])) & (x[best] * x[idx] + y[best] * y[idx] >= 0)]
        b2 = -1
        for i in tie:
            if b2 < 0: b2 = i; continue
            p, q = i, b2
            cr = (al[q] - L) * rho[p] - rho[q] * (al[p] - L)
            dp2 = (al[p] - L) ** 2 + rho[p] ** 2; dq2 = (al[q] - L) ** 2 + rho[q] ** 2
            t2 = TOL2 * MAXABS * (np.sqrt(dp2) + np.sqrt(dq2))
            if cr < -t2: b2 = p
            elif cr > t2: pass
            elif dp2 > dq2: b2 = p
        if b2 == best: break
        best = b2
    return b2

def hull(P):
    m = len(P)
    global MAXABS
    MAXABS = np.abs(P).max()
    tol = TOL1 * MAXABS
    cand = np.arange(m)
    for ax in range(3):
        lo = P[cand, ax].min()
        cand = cand[P[cand, ax] <= lo + (tol if ax < 2 else 0.0)]
    a = int(cand[0])
    A = P[a]; Vt = A + np.array([0, 0, 8 * np.abs(P).max() + 1.0])
    b = wrap(P, A, Vt, np.array([-1., 0, 0]), np.array([0, 1., 0]), (a, -1))
    B = P[b]
    eab = unit(B - A); rw = (Vt - A) - eab * ((Vt - A) @ eab)
    c = wrap(P, A, B, unit(np.cross(B - A, Vt - A)), unit(rw), (a, b))
    keep = {a, b, c}; done = {(b, a), (a, c), (c, b)}
    stack = [(b, a, c), (a, c, b), (c, b, a)]
    guard = 0
    while stack and guard < 8 * m + 64:
        guard += 1
        s, t, w = stack.pop()
        if (t, s) in done: continue
        Sp, Tp, Wp = P[s], P[t], P[w]
        e = unit(Tp - Sp); rw = (Wp - Sp) - e * ((Wp - Sp) @ e)
        c = wrap(P, Sp, Tp, unit(np.cross(Tp - Sp, Wp - Sp)), unit(rw), (s, t))
        if c < 0: continue
        keep.add(c); done |= {(t, s), (s, c), (c, t)}
        stack += [(s, c, t), (c, t, s)]
    return keep

if __name__ == '__main__':
    # usage: python profiles/tools/hull_wrap_np.py [tol1] [tol2] [closure iterations]
    # clusters: a lattice box surface (coplanar faces, collinear edges) under rotations of 0 ... 2 rad, random clouds, points
    # on a sphere; plus, when present, the contact clusters recorded from the oracle (gpurun_out/hull_cases.pkl)
    import os
    rng = np.random.default_rng(0)

    def rot(mag):
        w = rng.normal(size=3) * mag
        th = np.linalg.norm(w); k = w / th
        K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K

    g = np.linspace(-0.5, 0.5, 6)
    X, Y, Z = np.meshgrid(g, g * 0.4, g * 0.7, indexing='ij')
    box = np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1)
    surf = box[(np.abs(box[:, 0]) == 0.5) | (np.abs(box[:, 1]) == 0.2) | (np.abs(box[:, 2]) == 0.35)]
    base = [surf]
    if os.path.exists('gpurun_out/hull_cases.pkl'):
        base += [p for p, _ in pickle.load(open('gpurun_out/hull_cases.pkl', 'rb'))]
    cases = []
    for p in base:
        cases.append(p)
        for mag in (1e-15, 1e-12, 1e-9, 1e-6, 1e-3, 1.0, 2.0):
            cases.append(p @ rot(mag).T + rng.normal(size=3) * 0.2)
    for n in (5, 8, 20, 100):
        cases.append(rng.normal(size=(n, 3)))
    s_ = rng.normal(size=(200, 3)); cases.append(s_ / np.linalg.norm(s_, axis=1)[:, None])
    bad = 0
    for k, p in enumerate(cases):
        ref = {tuple(p[i]) for i in ConvexHull(p).vertices}
        got = {tuple(p[i]) for i in hull(p)}
        if got != ref:
            bad += 1
            print(k, len(p), 'extra', len(got - ref), 'missing', len(ref - got))
    print('cases', len(cases), 'different from Qhull', bad)
