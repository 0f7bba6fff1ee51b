"""Smallest case that touches every kernel (for compute-sanitizer): setup, logp, nuts, summary, map (both), predict."""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np
from fitoct_b200 import _abi as abi, _lib as L, synth
S = synth.make_profiles(3, modulated_only=True)
sel = np.arange(0, 481, 7)          # N = 69: partial last pass
b = abi.make_problems_dense(S["x"][sel], S["Y"][:, sel], S["UY"][:, sel], S["theta0"], S["Sigma0"], Nn=5, ids=S["ids"])
spec = abi.default_spec()
q = np.zeros((3, 2, 10)); q[:, :, :3] = S["theta0"][:, None, :]; q[:, :, 8] = np.log(0.1)
print("logp", L.logp_grad(0, b, 3, spec, q)[0].ravel()[:2])
out = L.sample(0, b, 3, spec, abi.default_cfg(n_warmup=12, n_iter=25, seed=2, chains=3, max_treedepth=5))
print("draws finite", np.isfinite(out["draws"][..., :10]).all(), "summary", out["summary"][0, 0, :3])
print("map", L.expgp_map(b, 3, spec)[2])
bm = abi.make_problems_dense(S["x"][sel], S["Y"][:, sel], S["UY"][:, sel], S["theta0"], S["Sigma0"], Nn=0)
print("mono", L.monoexp_map(bm, 3, abi.default_spec(1))[3])
print("predict", L.predict(0, b, 0, spec, out["draws"][0, :2, 0])[0].shape)
