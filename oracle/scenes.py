"""ORACLE (test infrastructure) -- instantiate a scene spec (diffsdfsim_b200/scenes.py) as an oracle World.

``params`` lets a test make selected scalars differentiable leaves:
``build(spec, params={'mass': t, 'fric_coeff': t, 'push': t2})`` apply to the LAST body (the moving object),
mirroring how the reference experiments parametrise their scenes
(experiments/system_identification/optim_sysid.py:105-131).
"""
import numpy as np
import torch

from diffsdfsim_b200 import meshes, scenes
from . import sdf as S
from .sim import Body, World, tens

F64 = torch.float64


def mesh_for(b):
    """(verts, faces) numpy buffers for a body spec -- meshes are hot-path inputs."""
    k = b['kind']
    if k == 'box':
        return meshes.box_mesh(b['dims'], b['max_tri_length'])
    if k == 'sphere':
        if isinstance(b['rad'], torch.Tensor):       # differentiable radius: unit icosphere scaled by it (bodies.py:1001-1009)
            v, f = meshes.icosphere(1.0, (b['mesh'] or {}).get('subdivisions', 4))
            return torch.as_tensor(v, dtype=F64) * b['rad'], f
        return meshes.icosphere(b['rad'], (b['mesh'] or {}).get('subdivisions', 4))
    if k == 'cylinder':
        return meshes.cylinder_mesh(b['rad'], b['height'], 32, b['max_tri_length'])
    if k == 'grid':
        return scenes.grid_mesh(b)
    if k in ('box_rounded', 'brick', 'bowl'):
        return scenes.extra_kind_mesh(b)
    raise ValueError(k)


def shape_for(b, mass):
    """(kind id, normalised params, scale, body-frame inertia) following bodies.py:778-1009 closed forms."""
    k = b['kind']
    if k == 'box':
        d = tens(b['dims'])
        sc = d.max() * 1.5 / 2
        I = mass * torch.diag(d[[1, 0, 0]] ** 2 + d[[2, 2, 1]] ** 2) / 12
        return S.BOX, [d / sc], sc, I
    if k == 'sphere':
        r = tens(b['rad'])
        sc = r * 1.5
        return S.SPHERE, [r / sc], sc, 2 / 5 * mass * r ** 2 * torch.eye(3, dtype=F64)
    if k == 'cylinder':
        r, h = tens(b['rad']), tens(b['height'])
        sc = torch.max(r, h / 2) * 1.5
        I = mass * torch.diag(torch.stack([(3 * r ** 2 + h ** 2) / 12, (3 * r ** 2 + h ** 2) / 12, r ** 2 / 2]))
        return S.CYLINDER, [r / sc, h / sc], sc, I
    if k in ('box_rounded', 'brick'):
        d, r = tens(b['dims']), tens(float(b['rad']))
        sc = d.max() * 1.5 / 2
        v, f = scenes.extra_kind_mesh(b)
        I = mass * tens(meshes.mesh_inertia(v, f))
        return (S.BOX_ROUNDED, [(d - 2 * r) / sc, r / sc], sc, I) if k == 'box_rounded' else (S.BRICK, [d / sc, r / sc], sc, I)
    if k == 'bowl':
        r, d = tens(float(b['rad'])), tens(float(b['height']))
        sc = (r + d) * 1.3333
        v, f = scenes.extra_kind_mesh(b)
        return S.BOWL, [r / sc, d / sc], sc, mass * tens(meshes.mesh_inertia(v, f))
    if k == 'grid':
        grid = tens(scenes.grid_array(b))
        return S.GRID, [grid], tens(b['scale']), mass * tens(scenes.grid_unit_inertia(b))
    raise ValueError(k)


def build(spec, params=None, max_iter=10, **world_kw):
    params = params or {}
    bodies, pinned = [], []
    n = len(spec['bodies'])
    for i, b in enumerate(spec['bodies']):
        last = i == n - 1
        if last and 'rad' in params:
            b = dict(b, rad=params['rad'])
        mass = params['mass'] if (last and 'mass' in params) else tens(float(b['mass']))
        if 'mass_all' in params:
            mass = params['mass_all'][i]
        fric = params['fric_coeff'] if ('fric_coeff' in params) else tens(float(b['fric_coeff']))
        kind, sp, scale, I = shape_for(b, mass)
        verts, faces = mesh_for(b)
        pos = params['pos'] if (last and 'pos' in params) else tens(b['pos'])
        vel = params['vel'] if (last and 'vel' in params) else tens(b['vel'])
        if 'vel_all' in params:
            vel = params['vel_all'][i]
        ob = Body(kind, sp, scale, verts if isinstance(verts, torch.Tensor) else torch.as_tensor(verts, dtype=F64),
                  torch.as_tensor(np.asarray(faces)).long(),
                  pos, vel, mass, I, b['restitution'], fric)
        if b['gravity']:
            ob.add_gravity()
        if b['ext_force'] is not None:
            f = tens(b['ext_force'])
            if last and 'push' in params:
                f = torch.cat([f[:3], params['push'][0:1], f[4:5], params['push'][1:2]])
            until = b['ext_until']
            ob.forces.append((lambda t, f=f, until=until: f if (until is None or t < until) else f * 0))
        bodies.append(ob)
        if b['pinned']:
            pinned.append(ob)
    for i, j in spec['no_contact']:
        bodies[i].add_no_contact(bodies[j])
    locks = [(bodies[i], a) for i, a in spec['axis_locks']]
    return World(bodies, pinned=pinned, axis_locks=locks, dt=spec['dt'], eps=spec['eps'], tol=spec['tol'],
                 fric_dirs=spec['fric_dirs'], strict_no_penetration=spec['strict_no_penetration'],
                 time_of_contact_diff=spec['time_of_contact_diff'], max_iter=max_iter,
                 post_stab=spec.get('post_stab', False), grippers=spec.get('grippers', ()), **world_kw)
