"""Small sampling workload for an ncu capture of the sampling kernel: N profiles, 60 + 40 iterations, summaries only.
usage: ncu_target.py [n_profiles] [Nn]   (prints the leapfrog count of the launch)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fitoct_b200 import _abi as abi, _lib as L, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 592
Nn = int(sys.argv[2]) if len(sys.argv) > 2 else 10
S = synth.make_profiles(n, modulated_only=True)
b = abi.make_problems_dense(S["x"], S["Y"], S["UY"], S["theta0"], S["Sigma0"], Nn=Nn, ids=S["ids"])
cfg = abi.default_cfg(n_warmup=60, n_iter=100, seed=1)
o = L.sample(0, b, n, abi.default_spec(), cfg, draws=False, summary=True)
print("profiles", n, "Nn", Nn, "leapfrogs", float(o["n_leapfrog"].sum()))
