"""ORACLE (test infrastructure, not product code) -- mixed-LCP primal-dual interior point.

Single-problem float64 restatement of the reference solver; a batch is solved
one problem at a time, which is exactly how the reference engine uses it
(nBatch == 1, lcp_physics/physics/engines.py:59-81) and what SURVEY.md s8
defines as parity for every whole-batch reduction in the reference.

Follows:
* one-time factorisation   lcp_physics/lcp/solvers/batch.py:413-479
* per-iteration factor     lcp_physics/lcp/solvers/batch.py:485-520
* KKT solve                lcp_physics/lcp/solvers/batch.py:380-410
* PDIPM loop / best iterate lcp_physics/lcp/solvers/batch.py:70-237
* implicit backward        lcp_physics/lcp/lcp.py:156-213

Linear algebra: the reference assembles a packed block-LU of
S = [[A Q^-1 A^T, A Q^-1 G^T],[G Q^-1 A^T, G Q^-1 G^T + F + D^-1]]; here the same
block elimination is applied with explicit (pivoted) LU factors of the two
diagonal blocks -- identical mathematics, round-off-level differences only
(pinned against the reference's own outputs in tests/golden).
"""
import torch

lu_factor = torch.linalg.lu_factor


def _lu_solve(fac, rhs):
    if rhs.dim() == 1:
        return torch.linalg.lu_solve(fac[0], fac[1], rhs.unsqueeze(1)).squeeze(1)
    return torch.linalg.lu_solve(fac[0], fac[1], rhs)


class Prefactor:
    """Everything that does not depend on the barrier scaling d."""

    def __init__(self, Q, G, A, F):
        self.neq = A.shape[0] if A is not None and A.numel() > 0 else 0
        self.Q_lu = lu_factor(Q)
        self.G, self.A = G, (A if self.neq else None)
        QinvGt = _lu_solve(self.Q_lu, G.t())
        self.R = G @ QinvGt + F
        if self.neq:
            QinvAt = _lu_solve(self.Q_lu, A.t())
            self.S11_lu = lu_factor(A @ QinvAt)
            self.S21 = G @ QinvAt                      # G Q^-1 A^T
            self.T = _lu_solve(self.S11_lu, self.S21.t())   # (A Q^-1 A^T)^-1 (A Q^-1 G^T)
            self.R = self.R - self.S21 @ self.T
        self.T22_lu = None

    def factor(self, d):
        self.T22_lu = lu_factor(self.R + torch.diag(1.0 / d))

    def solve(self, d, rx, rs, rz, ry):
        """Newton direction of the KKT system for residuals (rx, rs, rz, ry)."""
        G, A = self.G, self.A
        q = _lu_solve(self.Q_lu, rx)
        h2 = G @ q + rs / d - rz
        if self.neq:
            h1 = A @ q - ry
            y1 = _lu_solve(self.S11_lu, h1)
            w2 = -_lu_solve(self.T22_lu, h2 - self.S21 @ y1)
            w1 = -y1 - self.T @ w2
        else:
            w2 = -_lu_solve(self.T22_lu, h2)
            w1 = None
        g1 = -rx - G.t() @ w2
        if self.neq:
            g1 = g1 - A.t() @ w1
        dx = _lu_solve(self.Q_lu, g1)
        ds = (-rs - w2) / d
        return dx, ds, w2, w1


def _ratio_step(v, dv):
    """get_step: largest t with v + t dv >= 0 (IEEE semantics un-guarded, like the reference)."""
    a = -v / dv
    a[dv > 0] = max(1.0, a.max().item())
    return a.min()


def pdipm(Q, p, G, h, A, b, F, eps=1e-12, not_improved_lim=3, max_iter=20, return_trace=False):
    """Returns best-residual iterate (x, nu, lam, s) and the Prefactor (for the backward)."""
    nineq = G.shape[0]
    pf = Prefactor(Q, G, A, F)
    neq = pf.neq
    one = torch.ones(nineq, dtype=Q.dtype)
    zero_i = torch.zeros(nineq, dtype=Q.dtype)
    pf.factor(one)
    x, s, z, y = pf.solve(one, p, zero_i, -h, (-b if neq else None))
    if s.min() < 0:
        s = s - (s.min() - 1)
    if z.min() < 0:
        z = z - (z.min() - 1)

    best = None
    stalled = 0
    trace = []
    for it in range(max_iter):
        rx = G.t() @ z + Q @ x + p
        if neq:
            rx = rx + A.t() @ y
        rs = z
        rz = G @ x + s - h - F @ z
        ry = (A @ x - b) if neq else None
        mu = torch.abs((s * z).sum() / nineq)
        res = rz.norm() + rx.norm() + nineq * mu
        if neq:
            res = res + ry.norm()
        d = z / s
        try:
            pf.factor(d)
        except Exception:
            break
        if return_trace:
            trace.append(float(res))
        if best is None:
            best = [res, x.clone(), (y.clone() if neq else None), z.clone(), s.clone()]
            stalled = 0
        elif res < best[0]:
            best = [res, x.clone(), (y.clone() if neq else None), z.clone(), s.clone()]
            stalled = 0
        else:
            stalled += 1
        if stalled == not_improved_lim or best[0] < eps or mu > 1e32:
            break

        dxa, dsa, dza, dya = pf.solve(d, rx, rs, rz, ry)
        alpha = min(_ratio_step(z, dza), _ratio_step(s, dsa), torch.tensor(1.0, dtype=Q.dtype))
        sig = (((s + alpha * dsa) * (z + alpha * dza)).sum() / (s * z).sum()) ** 3
        rs2 = (-mu * sig + dsa * dza) / s
        dxc, dsc, dzc, dyc = pf.solve(d, torch.zeros_like(x), rs2, zero_i,
                                      (torch.zeros(neq, dtype=Q.dtype) if neq else None))
        dx, ds, dz = dxa + dxc, dsa + dsc, dza + dzc
        dy = (dya + dyc) if neq else None
        alpha = min(0.999 * min(_ratio_step(z, dz), _ratio_step(s, ds)), torch.tensor(1.0, dtype=Q.dtype))
        x = x + alpha * dx
        s = s + alpha * ds
        z = z + alpha * dz
        y = (y + alpha * dy) if neq else None
    out = (best[1], best[2], best[3], best[4], pf)
    return out + (trace,) if return_trace else out


def make_lcp_function(eps=1e-12, not_improved_lim=3, max_iter=20, check_Q_spd=True):
    """Factory mirroring ``LCPFunction(...)`` -> callable(Q,p,G,h,A,b,F) -> zhat (B,nz)."""

    class _Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, Q, p, G, h, A, b, F):
            B = Q.shape[0]
            outs, keep = [], []
            for i in range(B):
                if check_Q_spd and not torch.all(torch.linalg.eigvals(Q[i]).real > 0):
                    raise RuntimeError('Q is not SPD.')
                Ai = A[i] if A.numel() > 0 else None
                bi = b[i] if A.numel() > 0 else None
                x, nu, lam, s, pf = pdipm(Q[i], p[i], G[i], h[i], Ai, bi, F[i], eps, not_improved_lim, max_iter)
                outs.append(x)
                keep.append((nu, lam, s, pf))
            z = torch.stack(outs)
            ctx.keep = keep
            ctx.save_for_backward(z)
            ctx.has_eq = A.numel() > 0
            return z

        @staticmethod
        def backward(ctx, gz):
            (zhat,) = ctx.saved_tensors
            dQ, dp, dG, dh, dA, db, dF = [], [], [], [], [], [], []
            for i, (nu, lam, s, pf) in enumerate(ctx.keep):
                d = lam.clamp(min=1e-8) / s.clamp(min=1e-8)
                pf.factor(d)
                zi = torch.zeros_like(lam)
                dx, _, dlam, dnu = pf.solve(d, gz[i], zi, zi, (torch.zeros_like(nu) if pf.neq else None))
                z = zhat[i]
                dQ.append(0.5 * (torch.outer(dx, z) + torch.outer(z, dx)))
                dp.append(dx)
                dG.append(torch.outer(dlam, z) + torch.outer(lam, dx))
                dh.append(-dlam)
                dF.append(torch.outer(dlam, lam))
                if pf.neq:
                    dA.append(torch.outer(dnu, z) + torch.outer(nu, dx))
                    db.append(-dnu)
            st = torch.stack
            if ctx.has_eq:
                return st(dQ), st(dp), st(dG), st(dh), st(dA), st(db), st(dF)
            return st(dQ), st(dp), st(dG), st(dh), None, None, st(dF)

    def call(Q, p, G, h, A, b, F):
        return _Fn.apply(Q, p, G, h, A, b, F)

    call.Fn = _Fn
    return call
