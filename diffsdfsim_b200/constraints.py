"""Equality constraints (sdf_physics/physics3d/constraints.py:32-145): each supplies constant rows of Je.

All of these are ``static`` single-body constraints whose Jacobian is a 0/1 selection of that body's six
velocity components [wx,wy,wz,vx,vy,vz] -- represented here as a list of component indices, which is what the
dynamics kernel consumes (eq_rows (neq,2) = [body, component]).
"""


class _AxisConstraint:
    axes = ()

    def __init__(self, body1):
        self.static = True
        self.body1, self.body2 = body1, None
        self.num_constraints = len(self.axes)

    def rows(self):
        return list(self.axes)

    def J(self):
        import torch
        J = torch.zeros(len(self.axes), 6, dtype=torch.float64)
        for r, a in enumerate(self.axes):
            J[r, a] = 1
        return J, None

    def move(self, dt):
        pass

    def update_pos(self):
        pass


class XConstraint(_AxisConstraint):
    axes = (3,)


class YConstraint(_AxisConstraint):
    axes = (4,)


class ZConstraint(_AxisConstraint):
    axes = (5,)


class RotConstraint3D(_AxisConstraint):
    axes = (0, 1, 2)


class TotalConstraint3D(_AxisConstraint):
    """Pins a body: J = I6 (constraints.py:131-145 with lcp_physics/physics/constraints.py:212-214)."""
    axes = (0, 1, 2, 3, 4, 5)


class GripperJoint:
    """sdf_physics/physics3d/constraints.py:148-195: body2 may only move towards / away from body1 along ``axis`` (given in
    body1's frame) and must turn with it -- five equality rows coupling the two bodies, re-evaluated from the poses at
    every solve (``static = False``).  Not a 0/1 selection: a world with such a joint runs on the dense LCP operator and
    the host-driven step loop."""
    general = True

    def __init__(self, body1, body2, axis=(1.0, 0.0, 0.0)):
        self.static = False
        self.num_constraints = 5
        self.body1, self.body2 = body1, body2
        self.axis = [float(a) for a in axis]

    def rows(self):
        raise TypeError('GripperJoint has no selection rows')

    def J(self, p1=None, p2=None):
        """(J1, J2): (B,5,6) blocks on body1 / body2 for poses p1, p2 (B,7) (default: the bodies' current poses)."""
        import torch
        from .transforms import quaternion_apply
        p1 = self.body1.p if p1 is None else p1
        p2 = self.body2.p if p2 is None else p2
        p1 = p1 if p1.dim() == 2 else p1.unsqueeze(0)
        p2 = p2 if p2.dim() == 2 else p2.unsqueeze(0)
        B = max(p1.shape[0], p2.shape[0])
        p1, p2 = p1.expand(B, 7), p2.expand(B, 7)
        dt, dev = p1.dtype, p1.device
        eye = torch.eye(3, dtype=dt, device=dev)
        ax = quaternion_apply(p1[:, :4], torch.tensor(self.axis, dtype=dt, device=dev).expand(B, 3))
        k = ax.abs().argmin(dim=1)                                   # orthogonal(): e_k x ax, k = first argmin |ax_k|
        d1 = torch.linalg.cross(eye[k], ax)
        d2 = torch.linalg.cross(d1, ax)
        dirs = torch.nn.functional.normalize(torch.stack([d1, d2], 1), dim=2)          # (B,2,3)
        pos2 = p1[:, 4:] - p2[:, 4:]                                   # joint position (= body1's origin) seen from body2
        z = torch.zeros(B, dtype=dt, device=dev)
        skew2 = torch.stack([torch.stack([z, -pos2[:, 2], pos2[:, 1]], 1), torch.stack([pos2[:, 2], z, -pos2[:, 0]], 1),
                             torch.stack([-pos2[:, 1], pos2[:, 0], z], 1)], 1)          # (B,3,3)
        J1 = torch.zeros(B, 5, 6, dtype=dt, device=dev)
        J2 = torch.zeros(B, 5, 6, dtype=dt, device=dev)
        top1 = torch.cat([eye, torch.zeros(3, 3, dtype=dt, device=dev)], 1).expand(B, 3, 6)
        J1 = torch.cat([top1, torch.cat([torch.zeros(B, 2, 3, dtype=dt, device=dev), dirs], 2)], 1)
        J2 = torch.cat([-top1, torch.cat([dirs @ skew2, -dirs], 2)], 1)
        return J1, J2

    def move(self, dt):
        pass

    def update_pos(self):
        pass
