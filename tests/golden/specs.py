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
}
