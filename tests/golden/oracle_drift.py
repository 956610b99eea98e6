import sys, os, numpy as np, torch, warnings
warnings.filterwarnings('ignore')
sys.path[:0] = ['.', 'tests', 'tests/golden']
from specs import make_spec
from oracle.scenes import build
F64 = torch.float64
for name in sys.argv[1:]:
    g = np.load('tests/golden/%s.npz' % name)
    spec, leaves = make_spec(name, g)
    params = {k: torch.tensor(g['leaf_' + k], dtype=F64, requires_grad=True) for k in leaves}
    w = build(spec, params)
    loss = 0.; dp = []; dv = []; tries_ok = True
    for k in range(spec['steps']):
        w.step()
        tries_ok &= w.stats['substeps'][-1] == int(g['tries'][k])
        dp.append(float(np.abs(w.get_p().detach().numpy() - g['p'][k]).max())); dv.append(float(np.abs(w.v.detach().numpy() - g['v'][k]).max()))
        loss = loss + (w.bodies[-1].pos ** 2).sum()
    loss.backward()
    print(name, 'tries identical', tries_ok, 'loss rel', abs(float(loss) - float(g['loss'])) / abs(float(g['loss'])))
    print('  pose drift per step:', ' '.join('%.1e' % x for x in dp))
    print('  vel drift per step :', ' '.join('%.1e' % x for x in dv))
    for k, t in params.items():
        ref = g['grad_' + k]; print('  grad', k, 'max rel err %.2e' % (np.abs(t.grad.numpy() - ref).max() / max(np.abs(ref).max(), 1e-30)))
