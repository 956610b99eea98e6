"""Scene table shared by the golden generator (reference), the oracle tests and the CUDA parity tests."""
import numpy as np

from diffsdfsim_b200 import scenes

# name -> (spec factory, differentiable leaves {name: initial value})
SCENES = {
    'box_on_plane': (lambda: scenes.box_on_plane(floor=(4.0, 1.0, 4.0), steps=8),
                     dict(mass=1.0, fric_coeff=0.2, push=[3.0, 2.0])),
    # post-stabilisation on (engines.py:85-121, world.py:358-370; off by default in the reference)
    'box_on_plane_poststab': (lambda: dict(scenes.box_on_plane(floor=(4.0, 1.0, 4.0), steps=8), post_stab=True),
                              dict(mass=1.0, fric_coeff=0.2, push=[3.0, 2.0])),
    # a two-body joint with pose-dependent Je rows (GripperJoint, physics3d/constraints.py:148-195)
    'gripper_pair': (lambda: scenes.gripper_pair(steps=8), dict(mass=0.4, fric_coeff=0.2, push=[1.5, 0.8])),
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


# ---- contact lists for _filter_contacts (contacts.py:97-158): (name, p1 (n,3), normals (n,3))
def _rot(rng, mag):
    w = rng.normal(size=3) * mag
    th = np.linalg.norm(w)
    if th == 0:
        return np.eye(3)
    k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def filter_cases():
    rng = np.random.default_rng(0)
    up = np.array([0.0, 1.0, 0.0])
    cases = []
    # lattice on the surface of a box: 152 points, hull = 8 corners; every face coplanar, every edge collinear
    g = np.linspace(-0.5, 0.5, 6)
    X, Y, Z = np.meshgrid(g, g * 0.4, g * 0.7, indexing='ij')
    box = np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1)
    surf = box[(np.abs(box[:, 0]) == 0.5) | (np.abs(box[:, 1]) == 0.2) | (np.abs(box[:, 2]) == 0.35)]
    for mag in (0.0, 1e-6, 1e-3, 0.3, 1.0, 2.0):
        cases.append(('box lattice rot %g' % mag, surf @ _rot(rng, mag).T + rng.normal(size=3) * 0.2, None))
    # a resting rounded box: flat bottom patch + a chamfer ring slightly above it (the shape of the scene states)
    u = np.linspace(-0.27, 0.27, 13)
    bx, bz = np.meshgrid(u, u * 0.66, indexing='ij')
    bottom = np.stack([bx.ravel(), np.full(bx.size, -0.25), bz.ravel()], 1)
    ring = np.array([[sx * 0.31, -0.2494, z] for sx in (-1, 1) for z in np.linspace(-0.178, 0.178, 9)])
    for mag in (0.0, 1e-4, 1e-2):
        cases.append(('rounded box patch rot %g' % mag, np.concatenate([bottom, ring]) @ _rot(rng, mag).T, None))
    # random clouds (general position) and points on a sphere (all of them vertices)
    for n in (5, 9, 40, 300):
        cases.append(('cloud %d' % n, rng.normal(size=(n, 3)), None))
    s = rng.normal(size=(120, 3))
    cases.append(('sphere', s / np.linalg.norm(s, axis=1)[:, None], None))
    # planar (2-D path), with interior points and exact duplicates; collinear (1-D path); a single point twice
    pl = np.stack([rng.uniform(-1, 1, 80), np.zeros(80), rng.uniform(-1, 1, 80)], 1)
    cases.append(('planar', np.concatenate([pl, pl[:7]]), None))
    cases.append(('planar tilted', pl @ _rot(rng, 0.7).T + 0.3, None))
    t = rng.uniform(-1, 1, 30)
    cases.append(('collinear', np.stack([t, 0 * t, 2 * t], 1), None))
    cases.append(('short line', np.array([[0.0, 0, 0], [1e-4, 0, 0], [5e-4, 0, 0]]), None))
    # three normal clusters interleaved + zero normals (dropped)
    pts = rng.normal(size=(90, 3))
    nrm = np.zeros((90, 3))
    nrm[0::3] = [0, 1, 0]; nrm[1::3] = [1, 0, 0]; nrm[2::3] = _rot(rng, 0.004) @ up     # third cluster: 0.004 rad from the first
    nrm[5] = 0.0; nrm[17] = 0.0
    cases.append(('three clusters', pts, nrm))
    return [(name, p, (np.tile(up, (len(p), 1)) if n is None else n)) for name, p, n in cases]


def mesh_sdf_cases():
    """(name, sdf_func(points, *params), params, res) for the differentiable iso-surface mesh: an off-centre sphere (radius +
    centre) and an IGR-style decoder with its latent code as the parameter."""
    import torch
    from diffsdfsim_b200 import igr
    F64 = torch.float64
    dec = igr.init_decoder(seed=3, radius_init=0.6)

    def sphere(pts, r, c):
        return (pts - c).norm(dim=1) - r

    def decoder(pts, latent):
        return igr.decode(dec, latent, pts)
    return [('sphere', sphere, [torch.tensor(0.55, dtype=F64), torch.tensor([0.05, -0.02, 0.1], dtype=F64)], 24),
            ('decoder', decoder, [torch.tensor([0.12, -0.07], dtype=F64)], 24)]
