"""ev_sdf_utils stand-in (un-vendored CUDA extension, unpinned git HEAD)."""
from oracle.sdf import grid_interp  # noqa: F401


def marching_cubes(vol, iso):
    raise RuntimeError("marching_cubes is not available offline: hand meshes in as inputs (custom_mesh=True)")
