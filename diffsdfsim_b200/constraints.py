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
