import sys, json
sys.path.insert(0, "/root/repo")
import numpy as np
from fitoct_b200 import _abi as abi, _lib as L, synth
Nn = int(sys.argv[1]); n = int(sys.argv[2])
S = synth.make_profiles(n, modulated_only=True)
b = abi.make_problems_dense(S["x"], S["Y"], S["UY"], S["theta0"], S["Sigma0"], Nn=Nn, ids=S["ids"])
cfg = abi.default_cfg(n_warmup=500, n_iter=1500, seed=1)
plan = L.Plan(0, b, n, abi.default_spec(), cfg, want_draws=False, want_summary=True)
plan.run(1); plan.sync(); plan.run(2); plan.sync()
tm = plan.timing(); o = plan.fetch(); plan.close()
T = tm["sample_ms"] * 1e-3
print(sys.argv[3], "Nn", Nn, "grad/s %.4g" % (o["n_leapfrog"].sum() / T), "step_s %.3f" % T, tm["regs"], tm["blocks_per_sm"])
