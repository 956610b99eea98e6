"""Batched torch quaternion helpers for host-side conveniences (get_surface, scene set-up).

The stepping kernels carry their own device implementations (csrc/dsdf_math.cuh); these follow the same
pytorch3d.transforms conventions ((w,x,y,z), raw Hamilton products, 2/|q|^2 matrix scaling).
"""
import torch


def quaternion_raw_multiply(a, b):
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw], -1)


def quaternion_invert(q):
    return q * q.new_tensor([1.0, -1.0, -1.0, -1.0])


def quaternion_apply(q, p):
    shape = torch.broadcast_shapes(q.shape[:-1], p.shape[:-1])
    q = q.expand(*shape, 4)
    p = p.expand(*shape, 3)
    pq = torch.cat([p.new_zeros(*shape, 1), p], -1)
    return quaternion_raw_multiply(quaternion_raw_multiply(q, pq), quaternion_invert(q))[..., 1:]


def quaternion_to_matrix(q):
    r, i, j, k = q.unbind(-1)
    s = 2.0 / (q * q).sum(-1)
    m = torch.stack([1 - s * (j * j + k * k), s * (i * j - k * r), s * (i * k + j * r),
                     s * (i * j + k * r), 1 - s * (i * i + k * k), s * (j * k - i * r),
                     s * (i * k - j * r), s * (j * k + i * r), 1 - s * (i * i + j * j)], -1)
    return m.reshape(q.shape[:-1] + (3, 3))


def so3_exponential_map(w, eps=1e-4):
    """Rodrigues formula, |w|^2 clamped at eps before the sqrt; w (...,3) -> (...,3,3)."""
    nr = (w * w).sum(-1)
    th = torch.clamp(nr, eps).sqrt()
    inv = 1.0 / th
    f1 = inv * th.sin()
    f2 = inv * inv * (1.0 - th.cos())
    x, y, z = w.unbind(-1)
    o = torch.zeros_like(x)
    K = torch.stack([o, -z, y, z, o, -x, -y, x, o], -1).reshape(w.shape[:-1] + (3, 3))
    eye = torch.eye(3, dtype=w.dtype, device=w.device)
    return f1[..., None, None] * K + f2[..., None, None] * (K @ K) + eye
