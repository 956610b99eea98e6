"""Random-init IGR-style SDF decoders baked to grids (BASELINE config 4: no trained checkpoints exist offline).

The reference's "cow on pole" demo drops an IGR shape (demos/demo_meshsdf.py:121-142) whose decoder is IGR's
``ImplicitNet`` with the hyper-parameters of IGR_data/train_configs/bob_spot_setup.conf:8,39-45: latent size 2,
eight hidden layers of 128 units, a skip connection into layer 4, Softplus(beta = 100) and geometric initialisation.
This module restates that architecture (third-party: github.com/amosgropp/IGR, code/model/network.py -- not vendored
in the reference) with seeded random weights and samples it on the [-1,1]^3 lattice that ``SDFGrid3D`` consumes
(sdf_physics/physics3d/bodies.py:203-211: value at index (i,j,k) = f(x_i, y_j, z_k), linspace(-1,1,R) per axis).

Construction-time code (one batched GEMM chain per grid), not part of the stepping hot path.
"""
import math

import numpy as np
import torch


def init_decoder(seed, latent_size=2, hidden=(128,) * 8, skip_in=(4,), radius_init=1.0, d_pts=3, dtype=torch.float64):
    """Weights [(W_l, b_l)] of ImplicitNet(d_in = latent_size + 3) under geometric initialisation."""
    g = torch.Generator().manual_seed(int(seed))
    d_in = latent_size + d_pts
    dims = [d_in] + list(hidden) + [1]
    layers = []
    for l in range(len(dims) - 1):
        out_dim = dims[l + 1] - d_in if (l + 1) in skip_in else dims[l + 1]
        if l == len(dims) - 2:
            w = math.sqrt(math.pi) / math.sqrt(dims[l]) + 1e-5 * torch.randn(out_dim, dims[l], generator=g, dtype=dtype)
            b = torch.full((out_dim,), -float(radius_init), dtype=dtype)
        else:
            w = math.sqrt(2.0) / math.sqrt(out_dim) * torch.randn(out_dim, dims[l], generator=g, dtype=dtype)
            b = torch.zeros(out_dim, dtype=dtype)
        layers.append((w, b))
    return dict(layers=layers, skip_in=tuple(skip_in), beta=100.0, latent_size=latent_size)


def decode(dec, latent, pts):
    """f(latent, pts): pts (N,3), latent (latent_size,) -> (N,).  Input = cat([latent, pts]) like decode_igr
    (sdf_physics/physics3d/utils.py:330-336)."""
    dev, dt = pts.device, pts.dtype
    inp = torch.cat([latent.to(dev, dt).expand(pts.shape[0], -1), pts], 1)
    x = inp
    n = len(dec['layers'])
    for l, (w, b) in enumerate(dec['layers']):
        if l in dec['skip_in']:
            x = torch.cat([x, inp], -1) / math.sqrt(2.0)
        x = x @ w.to(dev, dt).T + b.to(dev, dt)
        if l < n - 1:
            x = torch.nn.functional.softplus(x, beta=dec['beta'])
    return x[:, 0]


def bake_grid(dec, latent, res=64, device='cpu', dtype=torch.float64, chunk=1 << 16):
    """(res,res,res) samples of the decoder on linspace(-1,1,res)^3, 'ij' order (what grid_sdf indexes)."""
    t = torch.linspace(-1.0, 1.0, res, dtype=dtype, device=device)
    pts = torch.stack(torch.meshgrid(t, t, t, indexing='ij'), 3).reshape(-1, 3)
    out = torch.empty(pts.shape[0], dtype=dtype, device=device)
    with torch.no_grad():
        for i in range(0, pts.shape[0], chunk):
            out[i:i + chunk] = decode(dec, latent, pts[i:i + chunk])
    return out.reshape(res, res, res)


def random_shape_grid(seed, res=64, radius_init=0.6, latent_scale=0.3, device='cpu', as_float32=True):
    """One seeded random-init decoder + latent -> baked grid (numpy float64; values rounded through float32 when
    ``as_float32`` so that a committed float32 fixture reproduces it exactly).

    radius_init: the configuration file says 1, which puts the zero level set of a fresh decoder ON the boundary of the
    sampling cube (|x| ~ 1) where marching cubes clips the surface and the central-difference field is zero
    (bodies.py:225-234); trained IGR shapes sit well inside the unit cube, so the stand-in uses 0.6 (documented in
    DESIGN.md)."""
    dec = init_decoder(seed, radius_init=radius_init)
    g = torch.Generator().manual_seed(int(seed) + 7919)
    latent = latent_scale * torch.randn(dec['latent_size'], generator=g, dtype=torch.float64)
    grid = bake_grid(dec, latent, res, device=device).cpu()
    if as_float32:
        grid = grid.to(torch.float32).to(torch.float64)
    return grid.numpy()
