#!/bin/bash
# round-2 GPU check F: parity suite, the full default bench line (run-until-converged, e2e, cpu baseline), ncu capture + launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/parity_metrics.jsonl
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
tail -8 gpurun_out/f_pytest.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --rhat-target 0 $BARGS > gpurun_out/f_bench_$name.json 2> gpurun_out/f_bench_$name.err; }
BARGS="" run slice4096_1000 A=1
BARGS="" run slice1024_1000 FOCT_SLICE_TICKS=1024
BARGS="" run noslice_1000 FOCT_SLICE_TICKS=0
BARGS="--profiles 1776" run slice4096_1776 A=1
BARGS="--profiles 1776" run noslice_1776 FOCT_SLICE_TICKS=0
BARGS="--profiles 2500" run slice4096_2500 A=1
BARGS="--profiles 2500" run noslice_2500 FOCT_SLICE_TICKS=0
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/f_bench_*slice*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("f_bench_")[1], "ms", round(d["ms_per_step"],1), "grad/s %.3e"%d["grad_per_s"], "frac %.3f"%d["roofline"]["frac"], "ess/s %.0f"%d["value"], d["roofline"]["launch"], "rhat_max %.3f"%d["quality"]["rhat_max"])
    except Exception as e:
        print(f, "failed", e, open(f.replace(".json",".err")).read()[-300:])
PY
timeout 900 python bench.py > gpurun_out/f_bench_default.json 2> gpurun_out/f_bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/f_bench_reference.json 2> gpurun_out/f_bench_reference.err; echo "ref rc=$?"
python scripts/ncu_target.py 888 > gpurun_out/f_target.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nuts2_kernel -c 1 -o gpurun_out/f_ncu_gb -f python scripts/ncu_target.py 888 > gpurun_out/f_ncu_gb.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/f_target.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/f_ncu_launch.log 2>&1
echo "launch list rc=$?"
python - <<'PY'
import json
for f in ("f_bench_default","f_bench_reference"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value","ms_per_step","grad_per_s","kernel_ms","quality","fixed_length","e2e","e2e_with_draws","gpu_launches")}, (d.get("roofline") or {}).get("frac"), d.get("cpu_baseline"))
    except Exception as e:
        print(f, "failed", e, open(f"gpurun_out/{f}.err").read()[-600:])
PY
