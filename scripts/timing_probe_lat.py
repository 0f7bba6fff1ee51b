"""Per-phase clock64 breakdown of a leaf for ONE profile: latency kernel against one chain per warp (-DFOCT_TIMING build)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fitoct_b200 import _abi as abi, _lib as L, synth
lib = L.lib()
def read(reset=1):
    a = (C.c_ulonglong * 8)()
    lib.foct_debug_timing(a, reset)
    return np.array(list(a), dtype=np.float64)
S = synth.make_profiles(2, modulated_only=True)
b = abi.make_problems_dense(S["x"], S["Y"][:1], S["UY"][:1], S["theta0"][:1], S["Sigma0"][:1], Nn=10)
cfg = abi.default_cfg(n_warmup=500, n_iter=1500, seed=1)
for mode in ("lat", "one"):
    if mode == "one": os.environ["FOCT_NO_LAT"] = "1"
    else: os.environ.pop("FOCT_NO_LAT", None)
    L.sample(0, b, 1, abi.default_spec(), cfg, draws=False, summary=False)
    read(1)
    t = time.perf_counter(); o = L.sample(0, b, 1, abi.default_spec(), cfg, draws=False, summary=False); dt = time.perf_counter() - t
    v = read(1)
    leaves, merges = v[6], v[7]
    names = ["grad total", "sweep loop", "reduce-scatter+exchange", "grad prologue", "leaf total", "merge loop"]
    print(f"{mode}: wall {dt:.3f}s leapfrogs {o['n_leapfrog'].sum():.0f} probe leaves {leaves:.3e} merges/leaf {merges/leaves:.2f}")
    for i, nm in enumerate(names):
        print(f"   {nm:24s} {v[i]/leaves:9.0f} cycles/leaf")
    print(f"   grad epilogue (priors)   {(v[0]-v[1]-v[2]-v[3])/leaves:9.0f}")
    print(f"   leaf non-grad non-merge  {(v[4]-v[0]-v[5])/leaves:9.0f}   (grad total includes init/init_stepsize evals)")
