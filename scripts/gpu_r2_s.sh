#!/bin/bash
# round-2 GPU check S: new defaults (software-pipelined basis loads + exp table in shared memory) through the parity and
# bit-identity tests; basis rows in shared memory (one CTA of 12 warps per SM, FOCT_BASIS_SMEM=1) against the L1 path
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_continue.py tests/test_gpu_parity.py -m gpu -q --timeout 240 --timeout-method thread > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s_pytest.log
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --rhat-target 0 $BARGS > gpurun_out/s_bench_$name.json 2> gpurun_out/s_bench_$name.err; }
BARGS="" run l1_1000 A=1
BARGS="--profiles 1776" run l1_1776 A=1
BARGS="" run bs_1000 FOCT_BASIS_SMEM=1
BARGS="--profiles 1776" run bs_1776 FOCT_BASIS_SMEM=1
BARGS="--profiles 3552" run bs_3552 FOCT_BASIS_SMEM=1
BARGS="--profiles 3552" run l1_3552 A=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/s_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("s_bench_")[1], "ms", round(d["ms_per_step"],1), "grad/s %.4e"%d["grad_per_s"], "frac %.4f"%d["roofline"]["frac"], "ess/s %.0f"%d["value"], d["roofline"]["launch"], "rhat_max %.6f"%d["quality"]["rhat_max"], "ess %.4f" % d["quality"]["mean_min_bulk_ess_per_profile"])
    except Exception as e:
        print(f, "failed", e, open(f.replace(".json",".err")).read()[-300:])
PY
