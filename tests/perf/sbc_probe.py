"""Simulation-based calibration (tests/test_gpu_scale.py::test_simulation_based_calibration) for several sampler seeds and
both sampling kernels: prints the chi-square p-value of every sampled quantity.  usage: sbc_probe.py [seed ...]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from scipy.stats import chi2 as chi2_dist
from fitoct_b200 import _abi as abi, _lib as L, synth

seeds = [int(a) for a in sys.argv[1:]] or [77, 78, 79]
rng = np.random.default_rng(20240607)
n, Nn = 640, 10
x = synth.depth_grid(); N = x.size
theta0 = np.array([1000.0, 2000.0, 300.0]); Sigma0 = np.diag((0.03 * theta0) ** 2); lam_rate = 60.0
uy = 0.5 * np.sqrt(2000.0 * np.exp(-x / 150.0) + 1.0)
spec = abi.default_spec()
one = abi.make_problems_dense(x, np.ones((1, N)), uy[None, :], theta0[None, :], Sigma0[None], Nn=Nn, lambda_rate=lam_rate)
B = L.basis(one, 0, spec)
th = theta0 + rng.standard_normal((n, 3)) * np.sqrt(np.diag(Sigma0))
lam = rng.gamma(2.0, 1.0 / lam_rate, n)
ygp = rng.standard_normal((n, Nn)) * lam[:, None]
sig = 1.0 + 0.1 * rng.standard_normal(n)
dL = ygp @ B
m = th[:, :1] + th[:, 1:2] * np.exp(-2.0 * x[None, :] / (th[:, 2:3] * (1.0 + dL)))
Y = m + sig[:, None] * uy[None, :] * rng.standard_normal((n, N))
truth = np.column_stack([th, ygp, lam, sig])
b = abi.make_problems_dense(x, Y, np.tile(uy, (n, 1)), np.tile(theta0, (n, 1)), np.tile(Sigma0, (n, 1, 1)), Nn=Nn, lambda_rate=lam_rate)
for kernel in ("one", "pair"):
    os.environ.pop("FOCT_FORCE_PAIR", None); os.environ.pop("FOCT_NO_PAIR", None)
    os.environ["FOCT_FORCE_PAIR" if kernel == "pair" else "FOCT_NO_PAIR"] = "1"
    for seed in seeds:
        for thin in (4, 8):
            cfg = abi.default_cfg(chains=4, n_warmup=400, n_iter=400 + 64 * thin, seed=seed)
            out = L.sample(abi.FOCT_EXPGP, b, n, spec, cfg, draws=True, summary=True)
            d = out["draws"][:, ::thin, :, :15]
            Ld = d.shape[1] * d.shape[2]
            ranks = (d.reshape(n, Ld, 15) < truth[:, None, :]).sum(axis=1)
            ps = []
            for k in range(15):
                h = np.bincount(np.minimum(ranks[:, k] * 16 // (Ld + 1), 15), minlength=16)
                ps.append(chi2_dist.sf(np.sum((h - n / 16) ** 2 / (n / 16)), 15))
            edge = np.mean((ranks < Ld // 16) | (ranks > Ld - Ld // 16), axis=0)
            print(kernel, "seed", seed, "thin", thin, "min p %.2e" % min(ps), "argmin", int(np.argmin(ps)), "ps", " ".join("%.2g" % p for p in ps),
                  "| edge max dev %.3f" % np.max(np.abs(edge - 0.125)), "div", out["n_divergent"].sum(), flush=True)
