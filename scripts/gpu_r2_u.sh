#!/bin/bash
# round-2 GPU check U: c x from the shared blob (8 KB staged per warp), low levels of the subtree stack in shared memory
# (A/B builds), the latency kernel (two warps per chain) on single profiles, parity and continuation tests of the full build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; lib=$2; shift 2; FOCT_LIB_PATH=$PWD/fitoct_b200/variants/lib_$lib.so timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --rhat-target 0 "$@" > gpurun_out/u_bench_$name.json 2> gpurun_out/u_bench_$name.err; }
for v in cx0 cx_s0 cx_s1 cx_s2 cx_s3; do
  run ${v}_1776 $v --profiles 1776
  run ${v}_1000 $v
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/u_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("u_bench_")[1], "ms", round(d["ms_per_step"],1), "grad/s %.4e"%d["grad_per_s"], "frac %.4f"%d["roofline"]["frac"], "ess/s %.0f"%d["value"], d["roofline"]["launch"], "rhat_max %.6f"%d["quality"]["rhat_max"], "ess %.4f" % d["quality"]["mean_min_bulk_ess_per_profile"])
    except Exception as e:
        print(f, "failed", e, open(f.replace(".json",".err")).read()[-300:])
PY
timeout 900 python -m pytest tests/test_gpu_continue.py tests/test_gpu_parity.py -m gpu -q -x --timeout 240 --timeout-method thread > gpurun_out/u_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/u_pytest.log
python - <<'PY'
import time, os, numpy as np
from fitoct_b200 import _abi as abi, _lib as L, synth
S = synth.make_profiles(5)
for Nn in (10, 15):
    b = abi.make_problems_dense(S["x"], S["Y"][1:2], S["UY"][1:2], S["theta0"][1:2], S["Sigma0"][1:2], Nn=Nn)
    cfg = abi.default_cfg(n_warmup=500, n_iter=1500, seed=1)
    for env in ("lat", "one"):
        if env == "one": os.environ["FOCT_NO_LAT"] = "1"
        else: os.environ.pop("FOCT_NO_LAT", None)
        L.sample(0, b, 1, abi.default_spec(), cfg)
        t = time.perf_counter(); o = L.sample(0, b, 1, abi.default_spec(), cfg); dt = time.perf_counter() - t
        print("single profile Nn", Nn, env, "wall_s %.4f" % dt, "leapfrogs", o["n_leapfrog"].sum(), "rhat_max %.4f" % o["summary"][0, :Nn + 5, 9].max(), "us/leapfrog/chain %.3f" % (dt * 1e6 / (o["n_leapfrog"].sum() / 4)))
PY
