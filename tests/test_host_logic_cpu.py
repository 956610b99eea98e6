"""Host-side logic that needs no GPU: trajectory time matching, the flat contact-set layout, analytic meshes."""
import numpy as np
import torch

F64 = torch.float64


def test_nearest_time_index_matches_the_reference_scan():
    """losses.nearest_time_index == the ascending scan of optim_sphere.py:121-139 (ties go to the LATER state)."""
    from diffsdfsim_b200.losses import nearest_time_index
    g = torch.Generator().manual_seed(0)
    for _ in range(20):
        S, Sp, W = 7, 9, 3
        t = torch.cumsum(torch.rand(S, W, generator=g, dtype=F64) * 0.05, 0)
        tt = torch.cumsum(torch.rand(Sp, W, generator=g, dtype=F64) * 0.05, 0)
        tt[2] = tt[1] + 0.0                      # duplicate target times: exact ties
        tt = torch.sort(tt, 0)[0]
        idx = nearest_time_index(t, tt)
        for w in range(W):
            last_j = 0
            for i in range(S):
                min_diff, last_diff, new_j = 1e100, 1e100, 0
                for j in range(last_j, Sp):
                    diff = abs(float(t[i, w]) - float(tt[j, w]))
                    if diff <= min_diff:
                        min_diff, new_j = diff, j
                    if diff > last_diff:
                        break
                    last_diff = diff
                assert int(idx[i, w]) == new_j
                last_j = new_j


def test_contact_set_flat_layout_clone_and_resize():
    from diffsdfsim_b200.contacts import ContactSet
    W, m = 5, 4
    cs = ContactSet(W, m, 'cpu')
    assert cs.count.shape == (W,) and cs.body.shape == (W, m, 2) and cs.abc.shape == (W, m, 3) and cs.geo.shape == (W, m, 10)
    for k in ('count', 'status', 'body', 'face', 'abc', 'geo'):
        t = getattr(cs, k)
        assert t.is_contiguous() and t.data_ptr() % 8 == 0
        assert t.untyped_storage().data_ptr() == cs.flat.untyped_storage().data_ptr()      # all views of one buffer
    cs.count[:] = torch.arange(W, dtype=torch.int32)
    cs.geo[:] = torch.arange(W * m * 10, dtype=F64).reshape(W, m, 10)
    c2 = cs.clone()
    c2.geo.zero_()
    assert float(cs.geo.sum()) > 0 and float(c2.geo.sum()) == 0 and torch.equal(c2.count, cs.count)
    big = cs.resized(2 * m)
    assert big.maxc == 2 * m and torch.equal(big.count, cs.count) and torch.equal(big.geo[:, :m], cs.geo)
    assert float(big.geo[:, m:].abs().sum()) == 0


def test_analytic_meshes_have_the_surveyed_sizes():
    """custom_mesh=True box meshes (bodies.py:799-854): the sizes SURVEY.md probed on the reference."""
    from diffsdfsim_b200 import meshes
    v, f = meshes.box_mesh([20.0, 1.0, 20.0], 0.1)
    assert v.shape == (89646, 3) and f.shape == (176000, 3)
    v, f = meshes.box_mesh([1.0, 1.0, 1.0], 0.1)
    assert v.shape == (726, 3) and f.shape == (1200, 3)
    assert np.abs(v).max() == 0.5 and f.min() == 0 and f.max() == 725
    # outward orientation: positive volume by the divergence theorem
    tri = v[f]
    vol = np.einsum('ij,ij->i', tri[:, 0], np.cross(tri[:, 1], tri[:, 2])).sum() / 6.0
    assert abs(vol - 1.0) < 1e-9
    v, f = meshes.icosphere(0.5, 4)
    assert v.shape == (2562, 3) and f.shape == (5120, 3)
    np.testing.assert_allclose(np.linalg.norm(v, axis=1), 0.5, rtol=1e-12)


def test_batched_trajectory_loss_matches_the_reference_function():
    """losses.trajectory_loss (batched, every world matched on its own) against the outputs of the reference's own
    trajectory_loss (experiments/trajectory_fitting/optim_sphere.py:114-160, executed by tests/golden/make_golden.py):
    irregular time stamps, exact ties in the nearest-time scan, gradients w.r.t. every recorded pose."""
    import os
    from types import SimpleNamespace
    import numpy as np
    import torch
    from diffsdfsim_b200.losses import trajectory_loss
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'trajectory_loss.npz'))
    W, S, St = g['t'].shape[0], g['t'].shape[1], g['tt'].shape[1]
    T = lambda a: torch.tensor(a, dtype=torch.float64)
    ps = [T(g['p'][:, k]).reshape(W, 2, 7).requires_grad_(True) for k in range(S)]
    a = SimpleNamespace(W=W, device=torch.device('cpu'), batched=True,
                        trajectory=[(T(g['t'][:, k]), ps[k], T(g['v'][:, k]).reshape(W, 2, 6), None, None) for k in range(S)])
    b = SimpleNamespace(W=W, device=torch.device('cpu'), batched=True,
                        trajectory=[(T(g['tt'][:, k]), T(g['pt'][:, k]).reshape(W, 2, 7), T(g['vt'][:, k]).reshape(W, 2, 6),
                                     None, None) for k in range(St)])
    loss = trajectory_loss(a, b)
    np.testing.assert_allclose(loss.detach().numpy(), g['loss'], rtol=1e-13)
    loss.sum().backward()
    grad = np.stack([x.grad.reshape(W, 14).numpy() for x in ps], 1)
    np.testing.assert_allclose(grad, g['grad'], rtol=1e-12, atol=1e-15)


def test_differentiable_iso_surface_mesh_and_inertia_match_the_reference_functions():
    """meshes.iso_surface_mesh (the reference's MeshSDF, bodies.py:652-704) and meshes.mesh_inertia_torch (get_ang_inertia,
    :260-395) against outputs of the reference's own functions (tests/golden/make_golden.py mesh_sdf; marching cubes =
    the declared stand-in): vertices, faces, inertia, and gradients w.r.t. a sphere's radius / centre and a decoder's
    latent code."""
    import os
    import numpy as np
    import torch
    from diffsdfsim_b200.meshes import iso_surface_mesh, mesh_inertia_torch
    from specs import mesh_sdf_cases
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'mesh_sdf.npz'))
    for name, fn, params, res in mesh_sdf_cases():
        ps = [q.clone().requires_grad_(True) for q in params]
        verts, faces = iso_surface_mesh(fn, ps, res=res)
        np.testing.assert_array_equal(faces.numpy(), g[name + '_faces'])
        np.testing.assert_allclose(verts.detach().numpy(), g[name + '_verts'], atol=1e-14, rtol=0)
        J = mesh_inertia_torch(verts * 2.0, faces, torch.tensor(1.3, dtype=torch.float64))
        np.testing.assert_allclose(J.detach().numpy(), g[name + '_J'], rtol=1e-10, atol=1e-13)
        loss = (torch.tensor(g[name + '_w']) * verts).sum() + (torch.tensor(g[name + '_u']) * J).sum()
        np.testing.assert_allclose(float(loss), float(g[name + '_loss']), rtol=1e-11)
        loss.backward()
        for k, q in enumerate(ps):
            np.testing.assert_allclose(q.grad.numpy(), g['%s_grad%d' % (name, k)], rtol=1e-9, atol=1e-12,
                                       err_msg='%s: gradient of parameter %d' % (name, k))


def test_neural_sdf_exact_query_matches_the_reference_query_sdfs():
    """bodies.SDFDecoder3D.query_sdfs(exact=True) -- the decoder evaluated like the reference's SDF3D.query_sdfs for bodies
    without a closed-form gradient (bodies.py:721-760) -- against outputs of the reference method itself
    (tests/golden/make_golden.py neural_query): values, directions, overlap mask, and the gradient of the point-cloud loss
    w.r.t. the latent code."""
    import os
    import numpy as np
    import torch
    from diffsdfsim_b200 import bodies, igr
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'neural_query.npz'))
    dec = igr.init_decoder(seed=3, radius_init=0.6)
    latent = torch.tensor([0.12, -0.07], dtype=torch.float64, requires_grad=True)
    body = bodies.SDFDecoder3D([0.1, 0.2, -0.1], 1.5, lambda pts, z: igr.decode(dec, z, pts), [latent], res=16)
    pts = torch.tensor(g['pts'])
    sd, gr, mask = body.query_sdfs(pts.clone(), return_grads=True, return_overlapmask=True, exact=True)
    np.testing.assert_array_equal(mask.numpy(), g['mask'])
    np.testing.assert_allclose(sd.detach().numpy(), g['sdf'], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(gr.detach().numpy(), g['dir'], rtol=1e-10, atol=1e-13)
    assert not sd.requires_grad                      # bodies.py:744-745: detached when the points are a leaf
    sd2, mask2 = body.query_sdfs(pts.clone(), return_grads=False, return_overlapmask=True, exact=True)
    loss = (torch.where(mask2, sd2, torch.zeros_like(sd2)) ** 2).sum()
    np.testing.assert_allclose(float(loss.detach()), float(g['loss']), rtol=1e-12)
    loss.backward()
    np.testing.assert_allclose(latent.grad.numpy(), g['glatent'], rtol=1e-10)
