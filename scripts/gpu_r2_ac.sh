#!/bin/bash
# round-2 GPU check AC: tail balancing of nuts2w_kernel (FOCT_TAIL_BALANCE=0 switches it off) + the scheduling bit-identity tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --rhat-target 0 $BARGS > gpurun_out/ac_bench_$name.json 2> gpurun_out/ac_bench_$name.err; }
BARGS="" run on_1000_a A=1
BARGS="" run off_1000_a FOCT_TAIL_BALANCE=0
BARGS="--seed 99" run on_1000_b A=1
BARGS="--seed 99" run off_1000_b FOCT_TAIL_BALANCE=0
BARGS="--profiles 1332" run on_1332 A=1
BARGS="--profiles 1332" run off_1332 FOCT_TAIL_BALANCE=0
BARGS="--profiles 1776" run on_1776 A=1
BARGS="--profiles 1776" run off_1776 FOCT_TAIL_BALANCE=0
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/ac_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("ac_bench_")[1], "ms", round(d["ms_per_step"],1), "grad/s %.4e"%d["grad_per_s"], "frac %.4f"%d["roofline"]["frac"], "rhat_max %.6f"%d["quality"]["rhat_max"], "ess %.4f" % d["quality"]["mean_min_bulk_ess_per_profile"])
    except Exception as e:
        print(f, "failed", e, open(f.replace(".json",".err")).read()[-300:])
PY
timeout 600 python -m pytest tests/test_gpu_continue.py tests/test_gpu_scale.py -m gpu -q -x --timeout 240 --timeout-method thread > gpurun_out/ac_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/ac_pytest.log
