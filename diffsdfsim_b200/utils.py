"""Defaults and small helpers (mirrors lcp_physics/physics/utils.py:33-67 and sdf_physics/physics3d/utils.py:41-62, 270-283)."""
import torch


class Defaults:
    DIM = 3
    EPSILON = 0.001          # contact detection threshold   (physics3d/utils.py:46)
    TOL = 1e-8               # penetration tolerance          (physics3d/utils.py:49)
    RESTITUTION = 0.5        # (lcp_physics/physics/utils.py:47)
    FRIC_COEFF = 0.9
    FRIC_DIRS = 8            # (physics3d/utils.py:52)
    FPS = 30
    DT = 1.0 / FPS
    ENGINE = 'PdipmEngine'
    CONTACT = 'FWContactHandler'
    SOLVER = 1
    DTYPE = torch.double
    DEVICE = torch.device('cuda:0')
    POST_STABILIZATION = False
    CUSTOM_MESH = True       # meshes are hot-path inputs here; marching-cubes meshing is a later row
    CUSTOM_INERTIA = True


Defaults3D = Defaults


def get_instance(mod, class_id):
    """lcp_physics/physics/utils.py:167-175: plug-in lookup by class-name string (or instantiate a class)."""
    if isinstance(class_id, str):
        return getattr(mod, class_id)()
    return class_id()


def get_tensor(x, base_tensor=None, **kw):
    """physics3d/utils.py:270-283."""
    if isinstance(x, torch.Tensor):
        return x
    if base_tensor is not None:
        return base_tensor.new_tensor(x, **kw)
    dev = Defaults.DEVICE if torch.cuda.is_available() else torch.device('cpu')
    return torch.tensor(x, dtype=Defaults.DTYPE, device=dev, **kw)


def default_device():
    return Defaults.DEVICE if torch.cuda.is_available() else torch.device('cpu')


def as_batched(x, trailing, device=None, dtype=torch.double):
    """Tensor with exactly one leading batch dim: accepts python scalars/lists, (trailing...) or (B, trailing...)."""
    t = x if isinstance(x, torch.Tensor) else torch.tensor(x, dtype=dtype)
    t = t.to(dtype=dtype)
    if device is not None:
        t = t.to(device)
    if t.dim() == trailing:
        t = t.unsqueeze(0)
    assert t.dim() == trailing + 1, 'expected %d or %d dims, got %s' % (trailing, trailing + 1, tuple(t.shape))
    return t
