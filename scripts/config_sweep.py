"""BASELINE.json configs[0..3] on one B200: MonoExp sampled, single ExpGP profile, Nn sweep 5..20 with uy from
foct_estimate_noise (the smoothing-spline residual law of the reference's estimateNoise, FitOCT.R:90-93), not a closed form."""
import json, sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fitoct_b200 import _abi as abi, _lib as L, synth
res = {}
S = synth.make_profiles(5)
# configs[0]: single profile, fitMonoExp sampled, 4 chains
b = abi.make_problems_dense(S["x"], S["Y"][:1], S["UY"][:1], S["theta0"][:1], S["Sigma0"][:1], Nn=0)
spec1 = abi.default_spec(abi.FOCT_MONOEXP)
th, _, _, _ = L.monoexp_map(b, 1, spec1)
cfg = abi.default_cfg(n_warmup=500, n_iter=1500, seed=1)
init = np.ascontiguousarray(np.tile(th[0], (4, 1))); cfg.init_mode = 2; cfg.init = abi.as_ptr(init)
L.sample(abi.FOCT_MONOEXP, b, 1, spec1, cfg)
t = time.perf_counter(); o = L.sample(abi.FOCT_MONOEXP, b, 1, spec1, cfg); dt = time.perf_counter() - t
res["c0_monoexp_single"] = dict(wall_s=dt, leapfrogs=float(o["n_leapfrog"].sum()), rhat_max=float(o["summary"][0, :3, 9].max()),
                                min_bulk_ess=float(o["summary"][0, :3, 10].min()))
# configs[1]: single ExpGP profile
b = abi.make_problems_dense(S["x"], S["Y"][1:2], S["UY"][1:2], S["theta0"][1:2], S["Sigma0"][1:2], Nn=10)
cfg = abi.default_cfg(n_warmup=500, n_iter=1500, seed=1)
L.sample(0, b, 1, abi.default_spec(), cfg)
t = time.perf_counter(); o = L.sample(0, b, 1, abi.default_spec(), cfg); dt = time.perf_counter() - t
res["c1_expgp_single"] = dict(wall_s=dt, leapfrogs=float(o["n_leapfrog"].sum()), rhat_max=float(o["summary"][0, :15, 9].max()),
                              min_bulk_ess=float(o["summary"][0, :15, 10].min()))
# configs[3]: Nn sweep with the estimateNoise-form uy
for Nn in (5, 10, 15, 20):
    n = 888 if Nn <= 15 else 592
    Sx = synth.make_profiles(n, modulated_only=True)
    b0 = abi.make_problems_dense(Sx["x"], Sx["Y"], Sx["UY"], Sx["theta0"], Sx["Sigma0"], Nn=Nn, ids=Sx["ids"])
    t = time.perf_counter(); en = L.estimate_noise(b0, n); t_noise = time.perf_counter() - t
    assert np.all(en["status"] <= 1)
    UY = np.stack(en["uy"])
    b = abi.make_problems_dense(Sx["x"], Sx["Y"], UY, Sx["theta0"], Sx["Sigma0"], Nn=Nn, ids=Sx["ids"])
    plan = L.Plan(0, b, n, abi.default_spec(), cfg, want_draws=False, want_summary=True)
    plan.run(1); plan.sync(); plan.run(2); plan.sync()
    tm = plan.timing(); o = plan.fetch(); plan.close()
    T = (tm["sample_ms"] + tm["summary_ms"]) * 1e-3
    leap = float(o["n_leapfrog"].sum())
    res[f"c3_Nn{Nn}"] = dict(profiles=n, step_s=T, grad_per_s=leap / T, tflops=leap * 481 * (4 * Nn + 23) / T / 1e12,
                             draws_per_s=n * 4000 / T, min_ess_per_s=float(np.nansum(np.nanmin(o["summary"][:, :Nn + 5, 10], axis=1))) / T,
                             rhat_q99=float(np.nanquantile(np.nanmax(o["summary"][:, :Nn + 5, 9], axis=1), 0.99)),
                             launch={k: tm[k] for k in ("grid", "blocks_per_sm", "regs", "smem_bytes")},
                             uy="foct_estimate_noise", estimate_noise_s=t_noise,
                             uy_over_true_sd_median=float(np.median(UY / Sx["UY"])), sigma_mean=float(np.nanmean(o["summary"][:, Nn + 4, 0])))
print(json.dumps(res))
