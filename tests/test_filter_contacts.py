"""``_filter_contacts`` (sdf_physics/physics3d/contacts.py:97-158) as a stand-alone device operator vs the oracle's
restatement of it (scipy's Qhull, like the reference): identical kept SETS on planar, linear, duplicate-ridden and
genuinely three-dimensional contact clusters, including the degenerate inputs Qhull sees in practice (lattice points on
the faces and edges of a box: coplanar facets, collinear edge points)."""
import os
import sys

import numpy as np
import pytest
import torch

from diffsdfsim_b200 import ops
from oracle.sim import hull_filter
from specs import filter_cases

F64 = torch.float64


def _load():
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'filter_contacts.npz'))
    cases = filter_cases()
    assert [c[0] for c in cases] == [str(s) for s in g['names']], 'regenerate tests/golden/filter_contacts.npz'
    return cases, g


@pytest.mark.gpu
def test_filter_matches_reference_and_oracle_sets():
    cases, gold = _load()
    K = max(len(p) for _, p, _ in cases)
    W = len(cases)
    P = torch.zeros(W, K, 3, dtype=F64); N = torch.zeros(W, K, 3, dtype=F64)
    cnt = torch.zeros(W, dtype=torch.int32)
    for w, (_, p, n) in enumerate(cases):
        P[w, :len(p)] = torch.as_tensor(p); N[w, :len(p)] = torch.as_tensor(n); cnt[w] = len(p)
    mask, status = ops.filter_contacts(N.cuda(), P.cuda(), eps=1e-3, count=cnt.cuda())
    mask, status = mask.cpu(), status.cpu()
    assert not bool((status & 4).any()), 'a cluster was too large for the device 3-D hull'
    for w, (name, p, n) in enumerate(cases):
        want = hull_filter(torch.as_tensor(n), torch.as_tensor(p), 1e-3)
        ref = gold['kept'][gold['off'][w]:gold['off'][w + 1]]            # the reference's own _filter_contacts
        np.testing.assert_array_equal(gold['p1'][w, :len(p)], p)
        assert {tuple(p[i]) for i in want.tolist()} == {tuple(p[i]) for i in ref.tolist()}, f'{name}: oracle vs reference'
        ref_pts = {tuple(p[i]) for i in ref.tolist()}
        got = mask[w, :len(p)].nonzero().flatten().tolist()
        got_pts = {tuple(p[i]) for i in got}
        # Qhull returns ONE of several exact duplicates; which index is not specified: compare the kept POINTS
        assert got_pts == ref_pts, f'{name}: kept points differ ({len(got_pts)} vs {len(ref_pts)})'
        assert len(got) == len(ref), f'{name}: kept count {len(got)} vs {len(ref)} (duplicates must be kept once)'


def test_oracle_filter_matches_reference_sets():
    """CPU: the oracle's restatement of _filter_contacts against the outputs of the reference function itself."""
    cases, gold = _load()
    for w, (name, p, n) in enumerate(cases):
        want = hull_filter(torch.as_tensor(n), torch.as_tensor(p), 1e-3)
        ref = gold['kept'][gold['off'][w]:gold['off'][w + 1]]
        assert {tuple(p[i]) for i in want.tolist()} == {tuple(p[i]) for i in ref.tolist()}, name
        assert len(want) == len(ref), name
