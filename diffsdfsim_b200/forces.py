"""External forces (sdf_physics/physics3d/forces.py:48-85, lcp_physics/physics/forces.py:38-80), batched.

A force returns a generalized force [torque(3); force(3)] of shape (B,6), B in {1, W}.
``force_func(t)`` receives the step's start time as a python float (the reference passes ``world.t``,
engines.py:36 / bodies.py:120-124); functions that must see each world's own sub-step time can set
``vectorized=True`` and will then be called with a (W,) tensor.
"""
import weakref

import torch

F64 = torch.float64


def constant_force(vec, until=None):
    """force_func returning ``vec`` ((6,) or (B,6)) for t < until (always if until is None)."""
    def fn(t):
        if until is None:
            return vec
        if isinstance(t, torch.Tensor):
            return vec * (t < until).to(vec.dtype).reshape(-1, 1)
        return vec if t < until else vec * 0
    return fn


def down_force(t):
    return ExternalForce3D.DOWN


class ExternalForce3D:
    UP = torch.tensor([0., 0, 0, 0, 1, 0], dtype=F64)
    DOWN = torch.tensor([0., 0, 0, 0, -1, 0], dtype=F64)
    RIGHT = torch.tensor([0., 0, 0, 1, 0, 0], dtype=F64)
    LEFT = torch.tensor([0., 0, 0, -1, 0, 0], dtype=F64)
    FRONT = torch.tensor([0., 0, 0, 0, 0, 1], dtype=F64)
    BACK = torch.tensor([0., 0, 0, 0, 0, -1], dtype=F64)
    ROTX = torch.tensor([1., 0, 0, 0, 0, 0], dtype=F64)
    ROTY = torch.tensor([0., 1, 0, 0, 0, 0], dtype=F64)
    ROTZ = torch.tensor([0., 0, 1, 0, 0, 0], dtype=F64)
    ZEROS = torch.zeros(6, dtype=F64)

    def __init__(self, force_func=down_force, multiplier=1., vectorized=False):
        self.force_func, self.multiplier, self.vectorized = force_func, multiplier, vectorized
        self.body = None

    def set_body(self, body):
        self.body = weakref.proxy(body)      # weak: no body <-> force cycle keeping an autograd graph alive

    def force(self, t):
        f = self.force_func(t)
        f = f.to(self.body.p.device) if isinstance(f, torch.Tensor) else torch.tensor(f, dtype=F64, device=self.body.p.device)
        if f.dim() == 1:
            f = f.unsqueeze(0)
        m = self.multiplier
        if isinstance(m, torch.Tensor) and m.dim() == 1:
            m = m.reshape(-1, 1)
        return f * m


class Gravity3D(ExternalForce3D):
    """Constant [0,0,0,0,-m g,0] (y up), physics3d/forces.py:73-85."""

    def __init__(self, g=10.0):
        self.multiplier, self.body, self.vectorized = g, None, False

    def force(self, t):
        down = ExternalForce3D.DOWN.to(self.body.p.device)
        return down.unsqueeze(0) * (self.body.mass * self.multiplier).reshape(-1, 1)
