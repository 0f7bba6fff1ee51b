"""Timing of the steps either side of the sampling path through the C ABI with host buffers,
beside the CPU oracle on a bounded sample.  Usage: python tests/perf/prep_bench.py [n_profiles] [n_cpu_sample]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fitoct_b200 import _abi as abi, _lib as L, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402  (CPU baseline leg only)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
n_cpu = int(sys.argv[2]) if len(sys.argv) > 2 else 100
S = synth.make_profiles(n)
eye = np.tile(np.eye(3), (n, 1, 1))
th00 = np.tile([0.0, 0.0, 1.0], (n, 1))
xy = abi.make_problems_dense(S["x"], S["Y"], np.ones_like(S["Y"]), th00, eye, Nn=0)
spec_m = abi.default_spec(abi.FOCT_MONOEXP)


def timed(f, reps=3):
    f()
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        r = f()
        ts.append(time.perf_counter() - t)
    return min(ts), r


t_noise, nz = timed(lambda: L.estimate_noise(xy, n, df=15.0))
UY = np.stack(nz["uy"])
mono = abi.make_problems_dense(S["x"], S["Y"], UY, th00, eye, Nn=0)
t_map, (th, H, br, st) = timed(lambda: L.monoexp_map(mono, n, spec_m))
t_gate, _ = timed(lambda: L.print_br(abi.FOCT_MONOEXP, mono, n, spec_m, br))
t_abc, _ = timed(lambda: L.estimate_exp_prior(mono, n, "abc", th, H))
print(f"GPU (C ABI, host buffers, incl. copies), {n} profiles x N={S['x'].size}:")
for name, t in (("estimateNoise", t_noise), ("fitMonoExp MAP", t_map), ("printBr gate", t_gate), ("estimateExpPrior abc", t_abc)):
    print(f"  {name:22s} {t * 1e3:9.2f} ms   {n / t:12.0f} profiles/s")

k = min(n, n_cpu)
sub = abi.make_problems_dense(S["x"], S["Y"][:k], np.ones_like(S["Y"][:k]), th00[:k], eye[:k], Nn=0)
t = time.perf_counter(); O.estimate_noise(sub, k, df=15.0); c_noise = time.perf_counter() - t
subm = abi.make_problems_dense(S["x"], S["Y"][:k], UY[:k], th00[:k], eye[:k], Nn=0)
t = time.perf_counter(); O.monoexp_map(subm, k, spec_m); c_map = time.perf_counter() - t
t = time.perf_counter(); O.exp_prior(subm, k, "abc", th[:k], H[:k]); c_abc = time.perf_counter() - t
print(f"CPU oracle, 1 thread, sample of {k} profiles:")
for name, t in (("estimateNoise", c_noise), ("fitMonoExp MAP", c_map), ("estimateExpPrior abc", c_abc)):
    print(f"  {name:22s} {t * 1e3:9.2f} ms   {k / t:12.0f} profiles/s")

if os.environ.get("PREP_PIPELINE", "1") == "1":
    from fitoct_b200 import api
    t = time.perf_counter()
    out = api.FitOCT_batch(S["x"], S["Y"], dict(nb_warmup=500, nb_sample=1000, Nn=10, priorType="abc"), chains=4)
    dt = time.perf_counter() - t
    s = out["expgp"]["summary"]
    print(f"foct_pipeline: {n} profiles, gate passed {out['n_expgp']} to fitExpGP (500+1000 iterations, 4 chains): {dt:.2f} s; "
          f"max split-Rhat {np.nanmax(s[:, :, 9]):.3f}, alerts by modulation kind "
          f"{[float(np.mean(out['alert'][S['mod_kind'] == m])) for m in range(5)]}")
