#!/bin/bash
# round-2 GPU check H: parity suite with the batch-size kernel choice, smoke, where the 4 % of the slicing restructure went, default bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/parity_metrics.jsonl
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h_pytest.log
tail -8 gpurun_out/h_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/h_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/h_smoke.log
V=$PWD/fitoct_b200/variants
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --rhat-target 0 $BARGS > gpurun_out/h_bench_$name.json 2> gpurun_out/h_bench_$name.err; }
BARGS="--profiles 1776" run new_1776 FOCT_SLICE_TICKS=0
BARGS="--profiles 1776" run old_1776 FOCT_LIB_PATH=$V/lib_old.so
BARGS="--profiles 1776" run nr_1776 FOCT_LIB_PATH=$V/lib_nr.so FOCT_SLICE_TICKS=0
BARGS="--profiles 1776" run ns_1776 FOCT_LIB_PATH=$V/lib_ns.so FOCT_SLICE_TICKS=0
BARGS="--profiles 1776" run nrs_1776 FOCT_LIB_PATH=$V/lib_nrs.so FOCT_SLICE_TICKS=0
BARGS="--profiles 600" run auto_600 A=1
BARGS="--profiles 600" run nopair_600 FOCT_NO_PAIR=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/h_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("h_bench_")[1], "ms", round(d["ms_per_step"],1), "grad/s %.3e"%d["grad_per_s"], "frac %.3f"%d["roofline"]["frac"], "ess/s %.0f"%d["value"], d["roofline"]["launch"], "rhat_max %.3f"%d["quality"]["rhat_max"])
    except Exception as e:
        print(f, "failed", e, open(f.replace(".json",".err")).read()[-300:])
PY
timeout 900 python bench.py > gpurun_out/h_bench_default.json 2> gpurun_out/h_bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/h_bench_default.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value","ms_per_step","grad_per_s","kernel_ms","quality","until_converged","e2e","e2e_with_draws","gpu_launches")}, (d.get("roofline") or {}).get("frac"), d.get("cpu_baseline"))
PY
