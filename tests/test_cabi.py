"""The C-ABI library loads and exports every symbol include/dsdf_b200.h declares (no compute calls: CPU-safe)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, 'include', 'dsdf_b200.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(dsdf_[a-z0-9_]+)\s*\(', txt)))


def test_library_builds_and_exports_header_symbols():
    from diffsdfsim_b200 import build, _lib
    path = build.build()
    assert os.path.exists(path)
    L = ctypes.CDLL(path)
    names = _declared()
    assert names, 'no declarations found in include/dsdf_b200.h'
    for n in names:
        assert hasattr(L, n), 'missing export: ' + n
    assert sorted(_lib.SIGNATURES) == names, 'python binding table out of sync with the header'
    assert _lib.lib().dsdf_version() >= 1


def test_smem_budget_of_the_box_on_plane_lcp():
    from diffsdfsim_b200 import _lib
    L = _lib.lib()
    assert L.dsdf_lcp_smem_bytes(12, 6, 100) < 227 * 1024
    assert L.dsdf_lcp_workspace_bytes(4, 12, 6, 100) == 4 * 100 * 100 * 8


def test_no_cpu_fallback():
    """Kernels refuse CPU tensors instead of silently computing elsewhere."""
    import pytest
    import torch
    from diffsdfsim_b200 import _lib
    from diffsdfsim_b200.lcp import LCPFunction
    z = torch.zeros(1, 2, dtype=torch.float64)
    with pytest.raises(_lib.DsdfLibraryError):
        LCPFunction()(torch.eye(2, dtype=torch.float64)[None], z, torch.ones(1, 1, 2, dtype=torch.float64),
                      torch.zeros(1, 1, dtype=torch.float64), torch.tensor([]), torch.tensor([]),
                      torch.zeros(1, 1, 1, dtype=torch.float64))
