"""B200-native differentiable stepping for DiffSDFSim: ``World3D.step()`` forward + backward for batches of worlds.

Host side mirroring ``sdf_physics.physics3d`` / ``lcp_physics`` (world, bodies, constraints, forces, engines, contacts,
lcp) over hand-written sm_100a kernels behind the C ABI of ``include/dsdf_b200.h`` (``libdsdf_b200.so``).  There is no
CPU fallback: the kernels need CUDA tensors, and a missing library raises ``DsdfLibraryError``.
"""
__version__ = '0.1.0'
__all__ = ['bodies', 'constraints', 'contacts', 'distributed', 'engines', 'forces', 'lcp', 'losses', 'meshes', 'ops',
           'scenes', 'transforms', 'utils', 'world']
