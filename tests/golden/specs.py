"""Scene table shared by the golden generator (reference), the oracle tests and the CUDA parity tests."""
from diffsdfsim_b200 import scenes

# name -> (spec factory, differentiable leaves {name: initial value})
SCENES = {
    'box_on_plane': (lambda: scenes.box_on_plane(floor=(4.0, 1.0, 4.0), steps=8),
                     dict(mass=1.0, fric_coeff=0.2, push=[3.0, 2.0])),
    'box_tilted': (lambda: scenes.box_on_plane(floor=(4.0, 1.0, 4.0), tilt=0.2, steps=10, floor_tri=0.2),
                   dict(mass=1.1, fric_coeff=0.15, push=[2.0, 4.0])),
    'bouncing_sphere': (lambda: scenes.bouncing_sphere(floor=(4.0, 1.0, 4.0), steps=14, floor_tri=0.2),
                        dict(fric_coeff=0.25, vel=[0.0, 0.0, 0.0, 2.0, 0.0, 0.0], pos=[0.0, 1.0, 0.0])),
    'grid_on_pole': (lambda: scenes.grid_on_pole(steps=8, with_floor=False), dict(mass=1.0, fric_coeff=0.15)),
    # config-3 shape: several free bodies + pinned floor, every pair searched (pins pair order and multi-pair contact rows)
    'mixed_primitives': (lambda: scenes.mixed_primitives(steps=8),
                         dict(mass=1.2, vel=[0.0, 0.0, 0.0, 0.2, -1.0, 0.1])),
    # radius fitting (the reference's published experiment, RESULTS.md:22-47): gradient of the rollout w.r.t. the sphere's
    # RADIUS, which enters through the mesh vertices, the SDF scale, and the inertia (optim_sphere.py:78-111)
    'sphere_radius': (lambda: scenes.bouncing_sphere(floor=(4.0, 1.0, 4.0), steps=14, floor_tri=0.2, subdivisions=3),
                      dict(rad=0.5, pos=[0.0, 1.0, 0.0])),
    # ---- BASELINE configurations at their named sizes
    # config 1: sphere dropped from 5 m with 5 m/s sideways on the 20x1x20 floor (176 000 faces), 100 steps
    # (experiments/trajectory_fitting/optim_sphere.py:78-111)
    'c1_bouncing_sphere': (lambda: scenes.bouncing_sphere(rad=1.0, height=5.0, vel=(0, 0, 0, 5.0, 0, 0), floor=(20.0, 1.0, 20.0),
                                                          steps=100, floor_tri=0.1, subdivisions=4),
                           dict(fric_coeff=0.25, vel=[0.0, 0.0, 0.0, 5.0, 0.0, 0.0], pos=[0.0, 5.0, 0.0])),
    # config 3 shape at size: 16 mixed primitives, every pair searched (nz = 96, no equality rows)
    'c3_mixed16': (lambda: scenes.mixed16(seed=0, steps=12, spacing=1.05, speed=2.0), dict(mass_all=None, vel_all=None)),
    # config 3, gravity + floor variant: 16 primitives resting on a pinned floor, ~100 simultaneous contacts (nz = 102)
    'c3_mixed16_floor': (lambda: scenes.mixed16_floor(seed=0, steps=6), dict(mass_all=None, vel_all=None)),
    # config 4 as the demo builds it: 64^3 grid baked from a random-init IGR-style decoder, 50x1x50 floor, scale 2, 33 steps
    'c4_cow_on_pole': (lambda grid=None: scenes.cow_on_pole(grid=grid, res=64, seed=1, steps=33),
                       dict(mass=1.0, fric_coeff=0.15, pos=[0.0, 6.0, 0.0])),
}


def default_leaves(spec, leaves):
    """Leaves given as None take the spec's own values (per-body arrays for the *_all parameters)."""
    out = dict(leaves)
    if out.get('mass_all', 0) is None:
        out['mass_all'] = [float(b['mass']) for b in spec['bodies']]
    if out.get('vel_all', 0) is None:
        out['vel_all'] = [list(b['vel']) for b in spec['bodies']]
    return out


def make_spec(name, g=None):
    """Spec of a golden scene; a grid committed with the golden file (float32, exact) replaces the bake recipe so that
    every machine sees bit-identical geometry."""
    mk, leaves = SCENES[name]
    if g is not None and 'grid_f32' in getattr(g, 'files', g):
        import numpy as np
        spec = mk(grid=np.asarray(g['grid_f32'], dtype=np.float64))
    else:
        spec = mk()
    return spec, default_leaves(spec, leaves)


def extra_sdf_kinds():
    """The analytic SDFs beyond box / sphere / cylinder (bodies.py:128-200): name -> (oracle kind id, normalised parameter
    tensors in the oracle's order, scale, kernel shape row [a,b,c,scale], kernel extra parameters).  Same bodies as the
    reference instances behind tests/golden/sdf_query.npz."""
    import torch
    from oracle import sdf as S
    F64 = torch.float64
    out = {}
    dims, r = torch.tensor([0.8, 0.5, 0.6], dtype=F64), 0.1
    sc = dims.max() * 1.5 / 2
    out['box_rounded'] = (S.BOX_ROUNDED, [(dims - 2 * r) / sc, torch.tensor(r, dtype=F64) / sc], sc,
                          torch.cat([(dims - 2 * r) / sc, sc.reshape(1)]), (float(r / sc), 0.0))
    dims = torch.tensor([1.0, 0.6, 0.5], dtype=F64)
    sc = dims.max() * 1.5 / 2
    out['brick'] = (S.BRICK, [dims / sc, torch.tensor(r, dtype=F64) / sc], sc, torch.cat([dims / sc, sc.reshape(1)]),
                    (float(r / sc), 0.0))
    rr, d = torch.tensor(0.6, dtype=F64), torch.tensor(0.15, dtype=F64)
    sc = (rr + d) * 1.3333
    out['bowl'] = (S.BOWL, [rr / sc, d / sc], sc, torch.stack([rr / sc, d / sc, torch.zeros((), dtype=F64), sc]), (0.0, 0.0))
    return out
