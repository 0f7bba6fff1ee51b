#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/parity_metrics.jsonl
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
tail -15 gpurun_out/d_pytest.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e $BARGS > gpurun_out/d_bench_$name.json 2> gpurun_out/d_bench_$name.err; }
V=$PWD/fitoct_b200/variants
BARGS="" run gb6_1000 A=1
BARGS="--profiles 1776" run gb6_1776 A=1
BARGS="" run gb5_1000 FOCT_LIB_PATH=$V/lib_gb5.so
BARGS="--profiles 1480" run gb5_1480 FOCT_LIB_PATH=$V/lib_gb5.so
BARGS="" run gb4u4_1000 FOCT_LIB_PATH=$V/lib_gb4u4.so
BARGS="--profiles 1184" run gb4u4_1184 FOCT_LIB_PATH=$V/lib_gb4u4.so
BARGS="" run nogb_1000 FOCT_NO_SHARED_BASIS=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/d_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("d_bench_")[1], "ms", round(d["ms_per_step"],1), "grad/s %.3e"%d["grad_per_s"], "frac %.3f"%d["roofline"]["frac"], "ess/s %.0f"%d["value"], d["roofline"]["launch"], "rhat_max %.3f"%d["quality"]["rhat_max"])
    except Exception as e:
        print(f, "failed", e, open(f.replace(".json",".err")).read()[-300:])
PY
