"""ev_sdf_utils stand-in (un-vendored CUDA extension, unpinned git HEAD)."""
import numpy as np
import torch

from oracle.sdf import grid_interp  # noqa: F401


def marching_cubes(vol, iso):
    """Declared stand-in (SURVEY.md s8c v): the iso-surface extraction of diffsdfsim_b200.meshes.surface_nets, returned the
    way the reference consumes it (bodies.py:664-667): vertices in INDEX coordinates, faces as a long tensor."""
    from diffsdfsim_b200.meshes import surface_nets
    res = vol.shape[0]
    v, f = surface_nets(vol.detach().cpu().numpy(), iso)
    verts = (torch.as_tensor(v, dtype=vol.dtype) + 1.0) / 2.0 * (res - 1)
    return verts, torch.as_tensor(f.astype(np.int64))
