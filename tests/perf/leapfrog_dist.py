"""Per-chain leapfrog counts of the bench workload (from the CPU oracle's timing build) -> npy, to model schedules."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from fitoct_b200 import _abi as abi, synth
from oracle import oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = synth.make_profiles(n)
b = abi.make_problems_dense(S["x"], S["Y"], S["UY"], S["theta0"], S["Sigma0"], Nn=10, ids=S["ids"])
cfg = abi.default_cfg(n_warmup=500, n_iter=1500, seed=1234, save_warmup=1)
t = time.time()
out = O.sample(abi.FOCT_EXPGP, b, n, abi.default_spec(), cfg, draws=True, summary=False, fast=True)
print("s", time.time() - t)
nl = out["sampler_params"][..., 3]  # [n, iter, chain]
np.save("/tmp/nleap.npy", nl)
print(nl.shape, nl.sum(axis=1).mean(), nl.sum(axis=1).min(), nl.sum(axis=1).max())
