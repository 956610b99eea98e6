"""Config-4 shaped contact detection for ncu: 256 worlds, per-world 64^3 grids + per-world iso-surface meshes, bodies
dropped from just above the pole so that the narrow phase (grid SDF queries against the pole / floor meshes and the body's
own faces against the analytic SDFs) is active within a few steps."""
import sys
sys.path.insert(0, '.')
import torch
from diffsdfsim_b200 import scenes
W = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
pw = scenes.per_world_grid_bodies(W, res=64, seed0=1000, scale=2.0, device='cuda')
spec = scenes.cow_on_pole(grid=pw['grid'][0].cpu().numpy(), steps=steps, drop=3.3)
# every body starts 5 cm above the pole top (y = 2): lowest vertex of its own mesh
idx = torch.arange(pw['verts'].shape[1], device='cuda')[None, :] < pw['nverts'][:, None]
miny = torch.where(idx, pw['verts'][:, :, 1], torch.full_like(pw['verts'][:, :, 1], 1e9)).min(1)[0]
pos = torch.zeros(W, 3, dtype=torch.float64, device='cuda')
pos[:, 1] = 2.05 - miny
world = scenes.build_world(spec, device='cuda', params=dict(pos=pos.clone().requires_grad_(True), **pw), strict_no_penetration=False)
for k in range(steps):
    world.step(fixed_dt=True)
    torch.cuda.synchronize()
    print(k, 'rounds', world.stats['rounds'][-1], 'contacts max', int(world.contact_set.count.max()), 'worlds with contacts', int((world.contact_set.count > 0).sum()), flush=True)
print('faces per world', int(pw['nfaces'].min()), int(pw['nfaces'].max()), 'stalled', world.stats.get('stalled', 0))
