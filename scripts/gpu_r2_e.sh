#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/parity_metrics.jsonl
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest.log
tail -12 gpurun_out/e_pytest.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e $BARGS > gpurun_out/e_bench_$name.json 2> gpurun_out/e_bench_$name.err; }
V=$PWD/fitoct_b200/variants
BARGS="" run lds128_1000 A=1
BARGS="--profiles 1184" run lds128_1184 A=1
BARGS="" run nolds128_1000 FOCT_LIB_PATH=$V/lib_nolds128.so
BARGS="--profiles 1184" run nolds128_1184 FOCT_LIB_PATH=$V/lib_nolds128.so
BARGS="--profiles 1776" run gb_lds128_1776 FOCT_SHARED_BASIS=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/e_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("e_bench_")[1], "ms", round(d["ms_per_step"],1), "grad/s %.3e"%d["grad_per_s"], "frac %.3f"%d["roofline"]["frac"], "ess/s %.0f"%d["value"], d["roofline"]["launch"], "rhat_max %.3f"%d["quality"]["rhat_max"])
    except Exception as e:
        print(f, "failed", e, open(f.replace(".json",".err")).read()[-300:])
PY
