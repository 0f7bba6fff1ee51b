"""First-contact GPU probe: parity of basis / logp / sampler vs the oracle, plus a first timing."""
import sys, time, json
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from fitoct_b200 import _abi as abi, _lib as L, synth
from oracle import oracle as O
np.set_printoptions(linewidth=200, precision=5, suppress=True)
print('devices', L.device_count())
print('fp64 peak', L.fp64_peak(0))
S = synth.make_profiles(8)
Nn = 10
b = abi.make_problems_dense(S['x'], S['Y'], S['UY'], S['theta0'], S['Sigma0'], Nn=Nn, ids=S['ids'])
spec = abi.default_spec()
# basis
Bg = L.basis(b, 1, spec); Bo = O.basis(b, 1, spec)
print('basis max abs diff', np.abs(Bg - Bo).max(), 'max|B|', np.abs(Bo).max())
# logp
rng = np.random.default_rng(0)
D = Nn + 5
q = np.zeros((8, 6, D))
for j in range(8):
    for k in range(6):
        q[j, k, :3] = S['theta0'][j] * (1 + 0.02 * rng.standard_normal(3))
        q[j, k, 3:3 + Nn] = 0.05 * rng.standard_normal(Nn)
        q[j, k, 3 + Nn] = np.log(0.1) + 0.3 * rng.standard_normal()
        q[j, k, 4 + Nn] = 0.2 * rng.standard_normal()
lp, g, c2 = L.logp_grad(0, b, 8, spec, q)
worst = 0
for j in range(8):
    Bj = L.basis(b, j, spec)
    lpo, go, c2o, at = O.logp_grad(0, b, j, spec, q[j], B=Bj, want_abs=True)
    e_lp = np.abs(lp[j] - lpo) / np.abs(lpo)
    e_g = np.abs(g[j] - go) / (np.abs(go) + at)
    e_c = np.abs(c2[j] - c2o) / np.abs(c2o)
    worst = max(worst, e_lp.max(), e_g.max(), e_c.max())
print('logp/grad worst rel err (shared basis):', worst)
# sampler short run: same seeds as oracle
cfg = abi.default_cfg(n_warmup=30, n_iter=60, seed=7, save_warmup=1)
t = time.time(); out = L.sample(0, b, 2, spec, cfg); print('gpu sample s', time.time() - t)
oo = O.sample(0, b, 2, spec, cfg)
for it in (0, 1, 2, 5, 10, 29, 30, 59):
    print(it, 'gpu', out['sampler_params'][0, it, 0], 'cpu', oo['sampler_params'][0, it, 0])
    print('   dq', np.abs(out['draws'][0, it, 0] - oo['draws'][0, it, 0]).max())
print('stepsize gpu', out['stepsize'][0], 'cpu', oo['stepsize'][0])
# full run timing
cfg = abi.default_cfg(n_warmup=500, n_iter=1500, seed=1234)
for n in (8,):
    t = time.time(); out = L.sample(0, b, n, spec, cfg, draws=True, summary=True); dt = time.time() - t
    nl = out['n_leapfrog'].sum()
    print(f'n={n} wall {dt:.2f}s leapfrogs {nl:.3e} -> {nl/dt:.3e} grad/s; div {out["n_divergent"].sum()}')
    so = O.summary(out['draws'][1])
    print('summary gpu vs oracle-on-gpu-draws max rel diff', np.nanmax(np.abs(out['summary'][1] - so) / (np.abs(so) + 1e-12)))
    print(out['summary'][1][:, [0, 2, 8, 9, 10]])
