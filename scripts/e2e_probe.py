"""Where does the end-to-end call spend its time?  FOCT_TRACE=1 python scripts/e2e_probe.py  (plan vs foct_sample, alternating)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from fitoct_b200 import _abi as abi, _lib as L  # noqa: E402

n = 1000
S, batch = bench.make_batch(n, 0, 10)
spec = abi.default_spec()
cfg = abi.default_cfg(n_warmup=500, n_iter=1500, seed=1234, chains=4)
plan = L.Plan(abi.FOCT_EXPGP, batch, n, spec, cfg, want_draws=False, want_summary=True)
for rnd in range(2):
    for s in range(2):
        plan.run(1237 + s); plan.sync(); tm = plan.timing()
        print(f"plan   seed {1237 + s}: sample {tm['sample_ms']:.1f} ms summary {tm['summary_ms']:.1f} ms", flush=True)
    for s in range(2):
        cfg.seed = 1237 + s
        t = time.perf_counter()
        L.sample(abi.FOCT_EXPGP, batch, n, spec, cfg, draws=False, summary=True)
        print(f"sample seed {1237 + s}: wall {1e3 * (time.perf_counter() - t):.1f} ms", flush=True)
plan.close()
