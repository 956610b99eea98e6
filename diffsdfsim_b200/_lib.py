"""ctypes binding of the C-ABI library (include/dsdf_b200.h).  No CPU fallback: a missing library is an error."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libdsdf_b200.so')
_lib = None

c_p = ctypes.c_void_p
c_i = ctypes.c_int
c_d = ctypes.c_double
c_sz = ctypes.c_size_t
c_ll = ctypes.c_longlong

# name -> (restype, argtypes).  Must list every symbol include/dsdf_b200.h declares (tests/test_cabi.py checks).
SIGNATURES = {
    'dsdf_version': (c_i, []),
    'dsdf_lcp_workspace_bytes': (c_sz, [c_i, c_i, c_i, c_i]),
    'dsdf_lcp_smem_bytes': (c_sz, [c_i, c_i, c_i]),
    'dsdf_lcp_forward': (c_i, [c_p] * 8 + [c_i] * 5 + [c_d, c_i, c_i, c_i] + [c_p] * 8),
    'dsdf_lcp_backward': (c_i, [c_p] * 10 + [c_i] * 5 + [c_p] * 10),
    'dsdf_sdf_query': (c_i, [c_i, c_p, c_p, c_i, c_ll, c_p, c_i, c_i, c_i, c_p, c_p, c_p]),
    'dsdf_sdf_query_backward': (c_i, [c_i, c_p, c_p, c_i, c_ll, c_p, c_i, c_i, c_p, c_p, c_p, c_p]),
    'dsdf_integrate': (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_p, c_p]),
    'dsdf_integrate_backward': (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_p, c_p, c_p, c_p, c_p]),
    'dsdf_contacts_workspace_bytes': (c_sz, [c_i, c_i, c_i]),
    'dsdf_contact_chunks_per_face_count': (c_i, [c_i]),
    'dsdf_contacts_detect': (c_i, [c_p, c_p, c_p, c_i, c_i, c_p, c_p, c_p, c_i, c_i, c_d, c_d, c_d, c_d, c_i, c_i, c_i]
                             + [c_p] * 10),
    'dsdf_contact_geometry_backward': (c_i, [c_p, c_p, c_p, c_i, c_i, c_d, c_i, c_i] + [c_p] * 7),
    'dsdf_dynamics_assemble': (c_i, [c_p] * 12 + [c_i] * 4 + [c_p] * 7),
    'dsdf_dynamics_assemble_backward': (c_i, [c_p] * 12 + [c_i] * 6 + [c_p] * 17),
}


class DsdfLibraryError(RuntimeError):
    pass


def lib():
    """Load (once) the CUDA library; raise loudly if it is absent -- there is no other compute path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DsdfLibraryError(
                'libdsdf_b200.so not found at %s: build it with `python -m diffsdfsim_b200.build` '
                '(there is no CPU / PyTorch fallback for the stepping kernels)' % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  Tensors must be contiguous."""
    if t is None:
        return None
    assert t.is_contiguous(), 'C ABI needs contiguous buffers'
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def check(rc, what):
    if rc != 0:
        if rc == -2:
            raise DsdfLibraryError('%s: problem too large for the shared-memory kernel' % what)
        raise DsdfLibraryError('%s failed with status %d' % (what, rc))


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise DsdfLibraryError('diffsdfsim_b200 kernels need CUDA tensors (got %s); there is no CPU fallback'
                                   % t.device)
