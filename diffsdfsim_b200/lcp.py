"""``LCPFunction`` -- drop-in for lcp_physics/lcp/lcp.py:43-214, backed by the sm_100a kernels.

    LCPFunction(eps=1e-12, verbose=0, notImprovedLim=3, max_iter=20, solver=1, check_Q_spd=True)
        -> callable(Q, p, G, h, A, b, F) -> zhat (nBatch, nz)

Same 7-in / 1-out signature, same broadcasting of un-batched inputs (their gradients are mean-reduced,
lcp.py:185-208), same exception types and messages.  One CTA per batch element; every whole-batch
reduction of the reference is per element here (SURVEY.md s8 a', s8).
"""
import torch

from . import _lib

F64 = torch.float64


def _expand(x, B, nd):
    """lcp_physics/lcp/util.py:95-101."""
    if x.dim() in (0, nd) or x.numel() == 0:
        return x, False
    if x.dim() == nd - 1:
        return x.unsqueeze(0).expand(B, *x.shape), True
    raise RuntimeError('Unexpected number of dimensions.')


def _batch(*xs):
    for x, nd in zip(xs, (3, 2, 3, 2, 3, 2, 3)):
        if x.dim() == nd:
            return x.shape[0]
    return 1


def lcp_solve_raw(Q, p, G, h, A, b, F, nineq_w=None, eps=1e-12, not_improved_lim=3, max_iter=20, check_spd=True,
                  nineq_smem=0):
    """Launch the forward kernel on contiguous f64 CUDA tensors; returns (x, nu, lam, s, status, iters)."""
    L = _lib.lib()
    _lib.require_cuda(Q, p, G, h, F)
    B, nz = p.shape
    ni = G.shape[1]
    neq = A.shape[1] if A is not None and A.numel() > 0 else 0
    dev = Q.device
    x = torch.empty(B, nz, dtype=F64, device=dev)
    nu = torch.empty(B, neq, dtype=F64, device=dev)
    lam = torch.empty(B, ni, dtype=F64, device=dev)
    s = torch.empty(B, ni, dtype=F64, device=dev)
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    iters = torch.zeros(B, dtype=torch.int32, device=dev)
    ws = torch.empty(L.dsdf_lcp_workspace_bytes(B, nz, neq, ni) // 8 + 1, dtype=F64, device=dev)
    rc = _lib.call('dsdf_lcp_forward', _lib.ptr(Q), _lib.ptr(p), _lib.ptr(G), _lib.ptr(h),
                            _lib.ptr(A) if neq else None, _lib.ptr(b) if neq else None, _lib.ptr(F),
                            _lib.ptr(nineq_w), B, nz, neq, ni, int(nineq_smem), eps, not_improved_lim, max_iter,
                            int(check_spd),
                            _lib.ptr(x), _lib.ptr(nu), _lib.ptr(lam), _lib.ptr(s), _lib.ptr(status), _lib.ptr(iters),
                            _lib.ptr(ws), _lib.stream())
    _lib.check(rc, 'dsdf_lcp_forward')
    return x, nu, lam, s, status, iters


def lcp_backward_raw(Q, G, A, F, x, nu, lam, s, gz, nineq_w=None, need=(True,) * 7, nineq_smem=0):
    L = _lib.lib()
    B, nz = x.shape
    ni = G.shape[1]
    neq = A.shape[1] if A is not None and A.numel() > 0 else 0
    dev = Q.device
    gz = gz.contiguous()
    mk = lambda flag, *shape: torch.empty(*shape, dtype=F64, device=dev) if flag else None
    dQ, dp, dG, dh = mk(need[0], B, nz, nz), mk(need[1], B, nz), mk(need[2], B, ni, nz), mk(need[3], B, ni)
    dA, db = mk(need[4] and neq, B, neq, nz), mk(need[5] and neq, B, neq)
    dF = mk(need[6], B, ni, ni)
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    ws = torch.empty(L.dsdf_lcp_workspace_bytes(B, nz, neq, ni) // 8 + 1, dtype=F64, device=dev)
    rc = _lib.call('dsdf_lcp_backward', _lib.ptr(Q), _lib.ptr(G), _lib.ptr(A) if neq else None, _lib.ptr(F), _lib.ptr(nineq_w),
                             _lib.ptr(x), _lib.ptr(nu), _lib.ptr(lam), _lib.ptr(s), _lib.ptr(gz),
                             B, nz, neq, ni, int(nineq_smem), _lib.ptr(dQ), _lib.ptr(dp), _lib.ptr(dG), _lib.ptr(dh), _lib.ptr(dA),
                             _lib.ptr(db), _lib.ptr(dF), _lib.ptr(status), _lib.ptr(ws), _lib.stream())
    _lib.check(rc, 'dsdf_lcp_backward')
    return dQ, dp, dG, dh, dA, db, dF


def LCPFunction(eps=1e-12, verbose=0, notImprovedLim=3, max_iter=20, solver=1, check_Q_spd=True):
    if solver != 1:
        raise NotImplementedError('only the batched PDIPM solver (solver=1) exists on this path')

    class LCPFunctionFn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, Q_, p_, G_, h_, A_, b_, F_):
            B = _batch(Q_, p_, G_, h_, A_, b_, F_)
            Q, _ = _expand(Q_, B, 3)
            p, _ = _expand(p_, B, 2)
            G, _ = _expand(G_, B, 3)
            h, _ = _expand(h_, B, 2)
            A, _ = _expand(A_, B, 3)
            b, _ = _expand(b_, B, 2)
            F, _ = _expand(F_, B, 3)
            neq = A.shape[1] if A.numel() > 0 else 0
            assert neq > 0 or G.shape[1] > 0
            c = lambda t: t.to(F64).contiguous()
            Q, p, G, h, F = c(Q), c(p), c(G), c(h), c(F)
            A, b = (c(A), c(b)) if neq else (None, None)
            x, nu, lam, s, status, _ = lcp_solve_raw(Q, p, G, h, A, b, F, None, eps, notImprovedLim, max_iter,
                                                     check_Q_spd)
            # one host sync for the error flags (the reference raises from inside forward too)
            st = int(torch.stack([(status & bit).max() for bit in (1, 2, 4, 8)]).sum().item()) if B else 0
            if st & 2:
                raise RuntimeError('Q is not SPD.')
            if st & 1:
                raise RuntimeError('\nqpth Error: Cannot perform LU factorization on Q.\n'
                                   'Please make sure that your Q matrix is PSD and has\na non-zero diagonal.\n')
            if (st & 8) and verbose >= 0:
                print('qpth warning: Returning an inaccurate and potentially incorrect solution.')
            ctx.save_for_backward(x, nu, lam, s, Q, G, A if neq else Q.new_empty(0), F)
            ctx.neq = neq
            ctx.shapes = [t.dim() for t in (Q_, p_, G_, h_, A_, b_, F_)]
            ctx.dtype = Q_.dtype
            return x.to(Q_.dtype)

        @staticmethod
        def backward(ctx, gz):
            x, nu, lam, s, Q, G, A, F = ctx.saved_tensors
            neq = ctx.neq
            need = list(ctx.needs_input_grad)
            grads = lcp_backward_raw(Q, G, A if neq else None, F, x, nu, lam, s, gz.to(F64), None, need)
            out = []
            for g, nd_in, nd_full in zip(grads, ctx.shapes, (3, 2, 3, 2, 3, 2, 3)):
                if g is None:
                    out.append(None)
                    continue
                if nd_in == nd_full - 1:
                    g = g.mean(0)
                out.append(g.to(ctx.dtype))
            return tuple(out)

    return LCPFunctionFn.apply
