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
    'dsdf_sdf_query_ex': (c_i, [c_i, c_p, c_d, c_d, c_p, c_i, c_ll, c_p, c_i, c_i, c_i, c_p, c_p, c_p]),
    'dsdf_sdf_query_backward_ex': (c_i, [c_i, c_p, c_d, c_d, c_p, c_i, c_ll, c_p, c_i, c_i, c_p, c_p, c_p, c_p]),
    'dsdf_integrate': (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_p, c_p]),
    'dsdf_integrate_backward': (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_p, c_p, c_p, c_p, c_p]),
    'dsdf_contacts_detect': (c_i, [c_p, c_p, c_i, c_p, c_p, c_p, c_i, c_i, c_d, c_d, c_d, c_d, c_i, c_i, c_i]
                             + [c_p] * 9),
    'dsdf_contacts_phase_cycles': (c_i, [c_p, c_i]),
    'dsdf_filter_contacts': (c_i, [c_p, c_p, c_p, c_i, c_i, c_d, c_p, c_p, c_p]),
    'dsdf_contact_geometry_backward': (c_i, [c_p, c_p, c_p, c_i, c_i, c_d, c_i, c_i] + [c_p] * 7),
    'dsdf_contact_geometry_backward_rows': (c_i, [c_p, c_p, c_p, c_i, c_i, c_d, c_i, c_i] + [c_p] * 8),
    'dsdf_contact_geometry_backward_full': (c_i, [c_p, c_p, c_p, c_i, c_i, c_d, c_i, c_i] + [c_p] * 10),
    'dsdf_dynamics_assemble': (c_i, [c_p] * 12 + [c_i] * 4 + [c_p] * 7),
    'dsdf_dynamics_assemble_backward': (c_i, [c_p] * 12 + [c_i] * 6 + [c_p] * 17),
    'dsdf_dynamics_solve_smem_bytes': (c_sz, [c_i] * 4),
    'dsdf_dynamics_solve': (c_i, [c_p] * 13 + [c_i] * 6 + [c_d, c_i, c_i] + [c_p] * 8),
    'dsdf_dynamics_solve_backward': (c_i, [c_p] * 13 + [c_i] * 8 + [c_p] * 14),
    'dsdf_toc_backward': (c_i, [c_i] * 3 + [c_p] * 9 + [c_d] + [c_p] * 7),
    'dsdf_contactset_move': (c_i, [c_i, c_i] + [c_p] * 16),
    'dsdf_attempt_commit': (c_i, [c_i] * 3 + [c_p] * 4 + [c_d, c_i, c_i] + [c_p] * 21),
    'dsdf_dynamics_solve_loop': (c_i, [c_p] * 13 + [c_i] * 6 + [c_d, c_i, c_i] + [c_p] * 9 + [c_i, c_i, c_p]),
    'dsdf_dynamics_solve_backward_loop': (c_i, [c_p] * 13 + [c_i] * 8 + [c_p] * 13 + [c_i, c_i, c_p]),
    'dsdf_contacts_detect_loop': (c_i, [c_p, c_p, c_i, c_p, c_p, c_p, c_i, c_i, c_d, c_d, c_d, c_d, c_i, c_i, c_i]
                                  + [c_p] * 11),
    'dsdf_dynamics_big_smem_bytes': (c_sz, [c_i] * 4),
    'dsdf_dynamics_big_workspace_bytes': (c_sz, [c_i] * 5),
    'dsdf_dynamics_big_solve': (c_i, [c_p] * 13 + [c_i] * 6 + [c_d, c_i, c_i] + [c_p] * 9 + [c_i, c_i, c_p, c_p]),
    'dsdf_dynamics_big_solve_backward': (c_i, [c_p] * 13 + [c_i] * 8 + [c_p] * 13 + [c_i, c_i, c_p, c_p]),
    'dsdf_step_begin': (c_i, [c_p, c_p]),
    'dsdf_step_resume': (c_i, [c_p, c_p]),
    'dsdf_step_rounds': (c_i, [c_p, c_i, c_i, c_i, c_p]),
    'dsdf_step_profile': (c_i, [c_i]),
    'dsdf_step_profile_read': (c_i, [c_p, c_p]),
    'dsdf_fma_peaks': (c_i, [c_i, c_p, c_p, c_p, c_p]),
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


# CUDA kernels launched by one call of each entry point (for bench.py's gpu_launches claim)
KERNELS_PER_CALL = {
    'dsdf_lcp_forward': 1, 'dsdf_lcp_backward': 1, 'dsdf_sdf_query': 1, 'dsdf_sdf_query_backward': 1, 'dsdf_sdf_query_ex': 1, 'dsdf_sdf_query_backward_ex': 1,
    'dsdf_integrate': 1, 'dsdf_integrate_backward': 1, 'dsdf_contacts_detect': 1,
    'dsdf_contact_geometry_backward': 1, 'dsdf_dynamics_assemble': 1, 'dsdf_dynamics_assemble_backward': 1,
    'dsdf_dynamics_solve': 1, 'dsdf_dynamics_solve_backward': 1, 'dsdf_attempt_commit': 1, 'dsdf_toc_backward': 1, 'dsdf_contactset_move': 1,
    'dsdf_step_begin': 1, 'dsdf_step_resume': 1, 'dsdf_step_rounds': 0, 'dsdf_dynamics_solve_backward_loop': 1, 'dsdf_dynamics_big_solve': 1, 'dsdf_dynamics_big_solve_backward': 1, 'dsdf_contact_geometry_backward_rows': 1, 'dsdf_contact_geometry_backward_full': 1,      # rounds: counted by the stepper (5 per round)
}
PROFILE = None      # set to {} to record (start, end) CUDA events around every entry-point call
LAUNCHES = {}       # name -> number of calls since reset_counters()


def reset_counters(profile=False):
    global PROFILE
    LAUNCHES.clear()
    PROFILE = {} if profile else None


def kernel_launches():
    return sum(KERNELS_PER_CALL.get(k, 1) * n for k, n in LAUNCHES.items())


def call(name, *args):
    """Invoke an entry point of the library (counts launches; optionally brackets it with CUDA events)."""
    fn = getattr(lib(), name)
    LAUNCHES[name] = LAUNCHES.get(name, 0) + 1
    if PROFILE is None:
        return fn(*args)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    rc = fn(*args)
    e.record()
    PROFILE.setdefault(name, []).append((s, e))
    return rc


def profile_summary():
    """name -> (calls, total ms) from the recorded events (synchronises)."""
    torch.cuda.synchronize()
    return {k: (len(v), sum(s.elapsed_time(e) for s, e in v)) for k, v in (PROFILE or {}).items()}


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
