#!/bin/bash
# round-2 GPU check AB: scales / priors of the latency kernel moved behind the sweep (A/B builds: lib_base = HEAD, lib_latp = the change)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in base latp base latp; do
FOCT_LIB_PATH=$PWD/fitoct_b200/variants/lib_$v.so python - "$v" <<'PY'
import sys, time, os, numpy as np
from fitoct_b200 import _abi as abi, _lib as L, synth
S = synth.make_profiles(5)
b = abi.make_problems_dense(S["x"], S["Y"][1:2], S["UY"][1:2], S["theta0"][1:2], S["Sigma0"][1:2], Nn=10)
cfg = abi.default_cfg(n_warmup=500, n_iter=1500, seed=1)
L.sample(0, b, 1, abi.default_spec(), cfg)
ts = []
for r in range(3):
    t = time.perf_counter(); o = L.sample(0, b, 1, abi.default_spec(), cfg); ts.append(time.perf_counter() - t)
print(sys.argv[1], "single profile wall_s", ["%.4f" % t for t in ts], "leapfrogs", o["n_leapfrog"].sum(), "us/leapfrog/chain %.3f" % (min(ts) * 1e6 / (o["n_leapfrog"].sum() / 4)), flush=True)
PY
done
run() { name=$1; lib=$2; shift 2; FOCT_LIB_PATH=$PWD/fitoct_b200/variants/lib_$lib.so timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --rhat-target 0 "$@" > gpurun_out/ab_bench_$name.json 2> gpurun_out/ab_bench_$name.err; }
run base_1776 base --profiles 1776
run latp_1776 latp --profiles 1776
run base_444 base --profiles 444
run latp_444 latp --profiles 444
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/ab_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("ab_bench_")[1], "ms", round(d["ms_per_step"],1), "grad/s %.4e"%d["grad_per_s"], "frac %.4f"%d["roofline"]["frac"], d["roofline"]["launch"], "rhat_max %.6f"%d["quality"]["rhat_max"])
    except Exception as e:
        print(f, "failed", e, open(f.replace(".json",".err")).read()[-300:])
PY
