"""CUDA LCP (dsdf_lcp_forward/backward through the C ABI) vs the oracle and vs reference golden instances."""
import os

import numpy as np
import pytest
import torch

from lcp_cases import contact_lcp
from oracle.lcp import make_lcp_function, pdipm
from specs import SCENES

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')
F64 = torch.float64


def _cuda(a):
    return torch.tensor(a, dtype=F64, device='cuda')


@pytest.mark.parametrize('name', list(SCENES))
def test_forward_matches_reference_instance(name):
    from diffsdfsim_b200.lcp import LCPFunction
    from diffsdfsim_b200 import _lib
    g = np.load(os.path.join(GOLD, name + '.npz'))
    if 'lcp_Q' not in g.files:
        pytest.skip('no LCP instance recorded for this scene')
    nz, ni, neq = g['lcp_G'].shape[2], g['lcp_G'].shape[1], g['lcp_A'].shape[1]
    if _lib.lib().dsdf_lcp_smem_bytes(nz, neq, ni) > 227 * 1024:
        pytest.skip('instance (%d inequality rows) exceeds the shared-memory budget of the dense one-CTA LCP kernel; '
                    'the fused structure-exploiting solver covers it (test_world_gpu golden rollout)' % ni)
    args = [_cuda(g['lcp_' + k]) for k in 'QpGhAbF']
    z = LCPFunction(max_iter=10, verbose=-1)(*args)
    np.testing.assert_allclose(z.cpu().numpy(), g['lcp_z'], rtol=1e-7, atol=1e-10)


@pytest.mark.parametrize('nc', [1, 2, 5, 10, 14])
def test_forward_backward_vs_oracle(nc):
    """Batch of contact-structured problems: solution, multipliers and all seven gradients."""
    from diffsdfsim_b200.lcp import LCPFunction, lcp_solve_raw
    rng = np.random.RandomState(nc)
    B = 6
    probs = [contact_lcp(rng, nc) for _ in range(B)]
    keys = 'QpGhAbF'
    cpu = [torch.tensor(np.stack([pr[k] for pr in probs]), dtype=F64, requires_grad=True) for k in keys]
    gpu = [t.detach().cuda().requires_grad_(True) for t in cpu]
    w = torch.tensor(rng.randn(B, probs[0]['Q'].shape[0]), dtype=F64)
    zo = make_lcp_function(max_iter=10)(*cpu)
    (zo * w).sum().backward()
    zg = LCPFunction(max_iter=10, verbose=-1)(*gpu)
    (zg * w.cuda()).sum().backward()
    np.testing.assert_allclose(zg.detach().cpu().numpy(), zo.detach().numpy(), rtol=1e-6, atol=1e-9)
    for k, a, b in zip(keys, gpu, cpu):
        ref = b.grad.numpy()
        np.testing.assert_allclose(a.grad.cpu().numpy(), ref, rtol=1e-4, atol=1e-6 * max(1.0, np.abs(ref).max()),
                                   err_msg='grad ' + k)
    # multipliers and slacks of the best iterate
    x, nu, lam, s, status, iters = lcp_solve_raw(*[t.detach() for t in gpu], max_iter=10)
    for i in range(B):
        xo, nuo, lamo, so, _ = pdipm(*[t.detach()[i] for t in cpu], max_iter=10)
        np.testing.assert_allclose(lam[i].cpu().numpy(), lamo.numpy(), rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(s[i].cpu().numpy(), so.numpy(), rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(nu[i].cpu().numpy(), nuo.numpy(), rtol=1e-5, atol=1e-9)
    assert int(status.max()) & 7 == 0


def test_ragged_contact_counts():
    """Per-world active row counts (padded to a common capacity), including a world with no contacts."""
    from diffsdfsim_b200.lcp import lcp_solve_raw
    rng = np.random.RandomState(7)
    counts = [0, 1, 3, 8, 8, 2]
    cap = 8
    probs = [contact_lcp(rng, max(c, 1), cap=cap) for c in counts]
    keys = 'QpGhAbF'
    gpu = [_cuda(np.stack([pr[k] for pr in probs])) for k in keys]
    nin = torch.tensor([10 * c for c in counts], dtype=torch.int32, device='cuda')
    x, nu, lam, s, status, iters = lcp_solve_raw(*gpu, nineq_w=nin, max_iter=10)
    for i, c in enumerate(counts):
        pr = probs[i]
        n = 10 * c
        T = lambda a: torch.tensor(a, dtype=F64)
        if c == 0:
            # equality-constrained QP: [[Q, A'],[A,0]] [x;y] = [-p; b]
            Q, A = pr['Q'], pr['A']
            K = np.block([[Q, A.T], [A, np.zeros((A.shape[0],) * 2)]])
            sol = np.linalg.solve(K, np.concatenate([-pr['p'], pr['b']]))
            np.testing.assert_allclose(x[i].cpu().numpy(), sol[:Q.shape[0]], rtol=1e-9, atol=1e-11)
            continue
        xo = pdipm(T(pr['Q']), T(pr['p']), T(pr['G'][:n]), T(pr['h'][:n]), T(pr['A']), T(pr['b']),
                   T(pr['F'][:n, :n]), max_iter=10)[0]
        np.testing.assert_allclose(x[i].cpu().numpy(), xo.numpy(), rtol=1e-6, atol=1e-9)
        assert float(lam[i, n:].abs().max() if n < 10 * cap else 0.0) == 0.0


def test_not_spd_raises_like_reference():
    from diffsdfsim_b200.lcp import LCPFunction
    rng = np.random.RandomState(0)
    pr = contact_lcp(rng, 2)
    pr['Q'] = pr['Q'].copy()
    pr['Q'][0, 0] = -1.0
    args = [_cuda(pr[k][None]) for k in 'QpGhAbF']
    with pytest.raises(RuntimeError, match='Q is not SPD'):
        LCPFunction(max_iter=10, verbose=-1)(*args)


def test_unbatched_inputs_broadcast_and_mean_reduce():
    from diffsdfsim_b200.lcp import LCPFunction
    rng = np.random.RandomState(3)
    pr = contact_lcp(rng, 3)
    Q = _cuda(pr['Q']).requires_grad_(True)                   # un-batched -> broadcast
    p = _cuda(np.stack([pr['p'], pr['p'] * 1.1])).requires_grad_(True)
    rest = [_cuda(pr[k][None].repeat(2, 0)) for k in 'GhAbF']
    z = LCPFunction(max_iter=10, verbose=-1)(Q, p, *rest)
    z.sum().backward()
    assert z.shape == (2, pr['Q'].shape[0]) and Q.grad.shape == Q.shape
    Qb = Q.detach()[None].repeat(2, 1, 1).requires_grad_(True)
    z2 = LCPFunction(max_iter=10, verbose=-1)(Qb, p.detach(), *rest)
    z2.sum().backward()
    np.testing.assert_allclose(Q.grad.cpu().numpy(), Qb.grad.mean(0).cpu().numpy(), rtol=1e-12, atol=1e-14)
