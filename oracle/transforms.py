"""ORACLE (test infrastructure, not product code) -- rotation helpers.

The reference calls ``pytorch3d.transforms`` (pinned 0.7.5 in
``environment.yaml:13``; source NOT under /root/reference).  This module
restates the *published* behaviour of the seven routines the stepping hot path
touches, in float64 torch so autograd supplies derivative oracles.  Call sites
in the reference that these stand in for:

* ``sdf_physics/physics3d/bodies.py:433,489,510,623,718``
* ``sdf_physics/physics3d/contacts.py:31-32,42,88,173-209``
* ``lcp_physics/physics/world.py:154-155,300-304``

Parity status: third-party, un-vendored -> "parity unpinned" for the routines
themselves; what IS pinned is that the reference's own code, executed here with
these routines injected as ``pytorch3d.transforms`` (tests/golden/ref_shims),
reproduces the committed golden vectors.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.
"""
import torch


def quaternion_raw_multiply(a, b):
    """Hamilton product, (w, x, y, z) layout, no sign fix-up."""
    aw, ax, ay, az = torch.unbind(a, -1)
    bw, bx, by, bz = torch.unbind(b, -1)
    return torch.stack((
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw), -1)


def standardize_quaternion(q):
    """Flip sign so that the real part is non-negative."""
    return torch.where(q[..., 0:1] < 0, -q, q)


def quaternion_multiply(a, b):
    """Hamilton product followed by standardisation (real part >= 0)."""
    return standardize_quaternion(quaternion_raw_multiply(a, b))


def quaternion_invert(q):
    """Conjugate (no normalisation)."""
    return q * q.new_tensor([1.0, -1.0, -1.0, -1.0])


def quaternion_apply(q, p):
    """Rotate points p (...,3) by q (...,4): (q * (0,p) * conj(q))[1:], raw products."""
    real = p.new_zeros(p.shape[:-1] + (1,))
    pq = torch.cat((real, p), -1)
    out = quaternion_raw_multiply(quaternion_raw_multiply(q, pq), quaternion_invert(q))
    return out[..., 1:]


def quaternion_to_matrix(q):
    """Rotation matrix of q with the 2/|q|^2 scaling (no pre-normalisation)."""
    r, i, j, k = torch.unbind(q, -1)
    s = 2.0 / (q * q).sum(-1)
    m = torch.stack((
        1 - s * (j * j + k * k), s * (i * j - k * r), s * (i * k + j * r),
        s * (i * j + k * r), 1 - s * (i * i + k * k), s * (j * k - i * r),
        s * (i * k - j * r), s * (j * k + i * r), 1 - s * (i * i + j * j)), -1)
    return m.reshape(q.shape[:-1] + (3, 3))


def _sqrt_positive_part(x):
    """sqrt(max(0,x)) with zero sub-gradient where x == 0."""
    ret = torch.zeros_like(x)
    pos = x > 0
    ret[pos] = torch.sqrt(x[pos])
    return ret


def matrix_to_quaternion(m):
    """Four-candidate conversion; pick the candidate with the largest |component|."""
    batch = m.shape[:-2]
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(m.reshape(batch + (9,)), -1)
    q_abs = _sqrt_positive_part(torch.stack((
        1.0 + m00 + m11 + m22,
        1.0 + m00 - m11 - m22,
        1.0 - m00 + m11 - m22,
        1.0 - m00 - m11 + m22), -1))
    cand = torch.stack((
        torch.stack((q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01), -1),
        torch.stack((m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20), -1),
        torch.stack((m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21), -1),
        torch.stack((m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2), -1)), -2)
    floor = torch.tensor(0.1, dtype=q_abs.dtype, device=q_abs.device)
    cand = cand / (2.0 * q_abs[..., None].max(floor))
    pick = torch.nn.functional.one_hot(q_abs.argmax(-1), num_classes=4) > 0.5
    return cand[pick, :].reshape(batch + (4,))


def hat(v):
    """Skew matrices of v (N,3)."""
    x, y, z = torch.unbind(v, -1)
    o = torch.zeros_like(x)
    return torch.stack((o, -z, y, z, o, -x, -y, x, o), -1).reshape(v.shape[:-1] + (3, 3))


def so3_exponential_map(log_rot, eps=1e-4):
    """Rodrigues formula with |w|^2 clamped at eps BEFORE the sqrt (N,3)->(N,3,3)."""
    nrms = (log_rot * log_rot).sum(-1)
    ang = torch.clamp(nrms, eps).sqrt()
    inv = 1.0 / ang
    f1 = inv * ang.sin()
    f2 = inv * inv * (1.0 - ang.cos())
    K = hat(log_rot)
    K2 = torch.bmm(K, K)
    return f1[:, None, None] * K + f2[:, None, None] * K2 + torch.eye(3, dtype=log_rot.dtype)[None]


def axis_angle_to_matrix(axis_angle):
    """Only needed so the reference's utils module imports under the shim."""
    ang = torch.norm(axis_angle, p=2, dim=-1, keepdim=True)
    half = ang * 0.5
    small = ang.abs() < 1e-6
    s = torch.empty_like(ang)
    s[~small] = torch.sin(half[~small]) / ang[~small]
    s[small] = 0.5 - (ang[small] * ang[small]) / 48
    return quaternion_to_matrix(torch.cat((torch.cos(half), axis_angle * s), -1))


def random_quaternions(n, dtype=None, device=None):
    o = torch.randn((n, 4), dtype=dtype, device=device)
    s = (o * o).sum(1)
    return o / torch.copysign(torch.sqrt(s), o[:, 0])[:, None]
