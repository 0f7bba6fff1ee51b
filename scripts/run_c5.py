"""BASELINE.json configs[4]: 100,000 synthetic profiles x 4 chains of fitExpGP (Nn=10, 500+1000 iterations), sharded
over all visible B200s by ONE call of the library (cfg.devices[], one host thread per GPU, no collective).
Prints one JSON line; summaries only (1e5 x 4 x 1000 x 17 doubles of draws = 54 GB stay on the devices)."""
import json, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from fitoct_b200 import _abi as abi, _lib as L, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
ndev = L.device_count()
t0 = time.perf_counter()
S = synth.make_profiles(n, modulated_only=True)
t_gen = time.perf_counter() - t0
b = abi.make_problems_dense(S["x"], S["Y"], S["UY"], S["theta0"], S["Sigma0"], Nn=10, ids=S["ids"])
cfg = abi.default_cfg(n_warmup=500, n_iter=1500, seed=1234)
t0 = time.perf_counter()
out = L.sample(abi.FOCT_EXPGP, b, n, abi.default_spec(), cfg, draws=False, summary=True, devices=list(range(ndev)))
wall = time.perf_counter() - t0
s = out["summary"]
rhat = np.nanmax(s[:, :15, 9], axis=1)
miness = np.nanmin(s[:, :15, 10], axis=1)
truth = np.array([1000.0, 2000.0, 300.0])
z = (s[:, :3, 0] - truth) / s[:, :3, 2]
print(json.dumps({
    "config": f"{n} profiles x 4 chains fitExpGP Nn=10, 500+1000 iterations, {ndev} x B200, one foct_sample() call",
    "wall_s": wall, "gen_s": t_gen, "profiles_per_s": n / wall, "draws_per_s": n * 4 * 1000 / wall,
    "min_ess_per_s": float(np.nansum(miness)) / wall, "grad_per_s": float(out["n_leapfrog"].sum()) / wall,
    "rhat_lt_1p01_frac": float(np.mean(rhat < 1.01)), "rhat_lt_1p05_frac": float(np.mean(rhat < 1.05)),
    "rhat_max": float(rhat.max()), "divergent_total": float(out["n_divergent"].sum()),
    "theta_within_3sd_frac": float(np.mean(np.abs(z) < 3)), "sigma_mean": float(s[:, 14, 0].mean()),
    "br_mean": float(s[:, 15, 0].mean()), "mean_min_bulk_ess": float(miness.mean())}))
