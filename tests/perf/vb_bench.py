"""Timing of method='vb' (ADVI) for a batch through the C ABI.  Usage: python tests/perf/vb_bench.py [n_profiles] [n_cpu]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fitoct_b200 import _abi as abi, _lib as L, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402  (CPU baseline leg only)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 16
S = synth.make_profiles(n, modulated_only=True)
batch = abi.make_problems_dense(S["x"], S["Y"], S["UY"], S["theta0"], S["Sigma0"], Nn=10)
spec = abi.default_spec(abi.FOCT_EXPGP)
cfg = abi.default_vb_cfg(omega0=-3.0)
L.vb(abi.FOCT_EXPGP, batch, min(n, 8), spec, cfg)
t = time.perf_counter(); g = L.vb(abi.FOCT_EXPGP, batch, n, spec, cfg); dt = time.perf_counter() - t
it = g["iters"].astype(float)
grads = float(np.sum(it * cfg.grad_samples + (it // cfg.eval_elbo) * cfg.elbo_samples + 6 * (cfg.adapt_iter + cfg.elbo_samples) + cfg.output_samples))
print(f"GPU foct_vb: {n} profiles (Nn=10, rstan::vb defaults, omega0=-3) in {dt:.3f} s = {n / dt:.0f} profiles/s; "
      f"converged {np.mean(g['status'] == 0):.3f}, median iterations {np.median(it):.0f}, ~{grads / dt:.3g} logp+grad sweeps/s")
sub = abi.make_problems_dense(S["x"], S["Y"][:k], S["UY"][:k], S["theta0"][:k], S["Sigma0"][:k], Nn=10)
t = time.perf_counter(); O.vb(abi.FOCT_EXPGP, sub, k, spec, cfg); dc = time.perf_counter() - t
print(f"CPU oracle, 1 thread: {k} profiles in {dc:.3f} s = {k / dc:.1f} profiles/s")
