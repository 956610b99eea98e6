"""SDF collision query (dsdf_contacts_detect through the C ABI) vs the oracle on states visited by oracle rollouts.

Parity protocol: (i) PRE-filter contact face-index sets bit-exact; (ii) the filtered contact set
equals the oracle's (scipy Qhull) set; (iii) contact geometry [n, p1, p2, pen] within 1e-9; (iv) pose gradients of the
geometry vs oracle autograd.
"""
import numpy as np
import pytest
import torch

from diffsdfsim_b200 import scenes
from oracle.scenes import build as build_oracle
from oracle.sim import contact_geometry as oracle_geometry
from specs import SCENES

pytestmark = pytest.mark.gpu
F64 = torch.float64
# scenes without a reference golden file: compared with the oracle only (whose filter calls scipy's Qhull like the
# reference).  The rounded box rests on a floor with its curved edges inside the contact band: its normal clusters are
# genuinely three-dimensional point sets, i.e. they exercise the 3-D convex hull (contacts.py:126-133).
EXTRA_SCENES = {'rounded_box': lambda: scenes.rounded_box_on_plane(steps=5)}


def _visited_states(name, max_states=48):
    """Run the oracle; record (pose vector, prefilter sets, final contacts) at every find_contacts call."""
    spec = EXTRA_SCENES[name]() if name in EXTRA_SCENES else SCENES[name][0]()
    w = build_oracle(spec)
    rec = []
    orig = w.find_contacts

    def spy():
        orig()
        rec.append((w.get_p().detach().clone(), [(a, b, ids.clone()) for a, b, ids in w.last_prefilter],
                    [(c[1], c[2], torch.cat([c[0][0], c[0][1], c[0][2], c[0][3].reshape(1)]).detach()) for c in w.contacts],
                    [(a, b, ids.clone(), abc.clone()) for a, b, ids, abc in w.last_search]))
    w.find_contacts = spy
    for _ in range(spec['steps']):
        w.step()
    # keep a spread of states, always including those with contacts
    with_c = [r for r in rec if r[2]]
    without = [r for r in rec if not r[2]]
    return spec, w, (with_c + without)[:max_states]


def _detector(spec, W, record=True, maxc=32):
    from diffsdfsim_b200.contacts import ContactDetector, GeometryTable
    bodies, _ = scenes.make_bodies(spec, device='cuda')
    table = GeometryTable(bodies, W, 'cuda')
    nb = len(bodies)
    pairs = [(i, j) for i in range(nb) for j in range(i + 1, nb) if bodies[j] not in bodies[i].no_contact]
    det = ContactDetector(table, pairs, W, nb, 'cuda', capK=768, maxc=maxc, record_prefilter=record)
    shape = torch.stack([b.shape_rows().expand(W, 4) for b in bodies], 1).contiguous()
    return bodies, table, det, shape, pairs


@pytest.mark.parametrize('name', ['box_on_plane', 'bouncing_sphere', 'grid_on_pole', 'box_tilted', 'rounded_box'])
def test_contact_sets_and_geometry_match_oracle(name):
    spec, ow, states = _visited_states(name)
    W = len(states)
    nb = len(spec['bodies'])
    bodies, table, det, shape, pairs = _detector(spec, W, maxc=320 if name in ('box_tilted', 'rounded_box') else 32)
    p = torch.stack([s[0] for s in states]).reshape(W, nb, 7).cuda().contiguous()
    cs = det.detect(p, shape, det.new_set(), eps=spec['eps'], tol=spec['tol'])
    torch.cuda.synchronize()
    cnt, pre_cnt = cs.count.cpu().numpy(), cs.pre_cnt.cpu().numpy()
    pre_ids, body, geo = cs.pre_ids.cpu().numpy(), cs.body.cpu().numpy(), cs.geo.cpu().numpy()
    status = cs.status.cpu().numpy()
    assert not np.any(status & 1), 'candidate capacity overflow'
    # contact-capacity overflow is only acceptable on penetrating (to-be-rejected) attempts, which keep ALL contacts
    assert not np.any((status & 2) & ~((status & 8) >> 2)), 'contact capacity overflow on a valid state'
    n_same_filtered = n_cmp = 0
    for w in range(W):
        _, pre, final, _ = states[w]
        # (i) pre-filter sets, per searched direction, bit-exact
        got = {}
        for d in range(2 * len(pairs)):
            if pre_cnt[w, d] >= 0:
                i, j = pairs[d // 2]
                key = (j, i) if d % 2 else (i, j)
                got[key] = sorted(pre_ids[w, d, :pre_cnt[w, d]].tolist())
        want = {(a, b): sorted(ids.tolist()) for a, b, ids in pre}
        assert {k: v for k, v in got.items() if v or k in want} == {k: want.get(k, []) for k in got if got[k] or k in want}, \
            f'state {w}: pre-filter contact face sets differ'
        # (ii)+(iii) filtered contacts: compare as sets keyed by (bodies, rounded p1)
        if status[w] & 2:
            continue
        n_cmp += 1
        mine = [(int(body[w, k, 0]), int(body[w, k, 1]), geo[w, k]) for k in range(cnt[w])]
        if len(mine) == len(final):
            used, ok = set(), True
            for (a, b, g) in mine:
                best = None
                for t, (fa, fb, fg) in enumerate(final):
                    if t in used or (fa, fb) != (a, b):
                        continue
                    err = np.abs(fg.numpy() - g).max()
                    if best is None or err < best[0]:
                        best = (err, t)
                if best is None or best[0] > 1e-7:
                    ok = False
                    break
                used.add(best[1])
            n_same_filtered += ok
    # the device hull filter must reproduce Qhull's vertex choice (scipy on the host in the reference) on EVERY state
    print(f'{name}: filtered contact sets identical on {n_same_filtered}/{n_cmp} states ({W} states pre-filter exact)')
    assert n_same_filtered == n_cmp, f'{name}: {n_cmp - n_same_filtered} of {n_cmp} filtered contact sets differ from Qhull'
    assert not np.any(status & 4), 'a non-planar cluster fell back to keep-all (DSDF_CON_HULL3D)'


@pytest.mark.parametrize('name', ['box_on_plane', 'bouncing_sphere', 'grid_on_pole'])
def test_geometry_pose_gradients_match_oracle_autograd(name):
    from diffsdfsim_b200.contacts import differentiable_geometry
    spec, ow, states = _visited_states(name)
    states = [s for s in states if s[2]][:8]
    W = len(states)
    nb = len(spec['bodies'])
    bodies, table, det, shape, pairs = _detector(spec, W)
    p = torch.stack([s[0] for s in states]).reshape(W, nb, 7).cuda().contiguous().requires_grad_(True)
    cs = det.detect(p.detach(), shape, det.new_set(), eps=spec['eps'], tol=spec['tol'])
    geo = differentiable_geometry(p, shape, cs, table)
    gen = torch.Generator().manual_seed(0)
    wgt = torch.randn(geo.shape, generator=gen, dtype=F64).cuda()
    mask = (torch.arange(geo.shape[1], device='cuda')[None, :] < cs.count[:, None]).unsqueeze(-1)
    (geo * wgt * mask).sum().backward()
    gp = p.grad.cpu().numpy()
    cnt, body, face, abc = cs.count.cpu(), cs.body.cpu(), cs.face.cpu(), cs.abc.cpu()
    wc = wgt.cpu()
    for w in range(W):
        po = states[w][0].clone().requires_grad_(True)
        ow.set_p(po)
        tot = 0.
        for k in range(int(cnt[w])):
            b1, b2 = ow.bodies[int(body[w, k, 0])], ow.bodies[int(body[w, k, 1])]
            n, p1, p2, pen = oracle_geometry(b1, b2, abc[w, k:k + 1], face[w, k:k + 1].long())
            tot = tot + (torch.cat([n[0], p1[0], p2[0], pen]) * wc[w, k]).sum()
        tot.backward()
        ref = po.grad.numpy().reshape(nb, 7)
        # NB flat-on-flat (box_on_plane): both finite-difference Laplacians vanish, so whether the normal comes from
        # b2 or from -b1 (contacts.py:199-202) is decided by round-off in the reference itself.  The contact kernels
        # are compiled without FMA contraction precisely so that such ties break as in torch's un-fused arithmetic.
        np.testing.assert_allclose(gp[w], ref, rtol=1e-8, atol=1e-8 * max(1.0, np.abs(ref).max()))
