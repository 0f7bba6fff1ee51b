#!/bin/bash
# round-2 GPU check N: the kernel as shipped — parity suite, slice length, the default bench line, reference arm, ncu capture, launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/parity_metrics.jsonl
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/n_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/n_smoke.log
timeout 900 python -m pytest tests -m gpu -q --timeout 240 --timeout-method thread > gpurun_out/n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n_pytest.log
tail -8 gpurun_out/n_pytest.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --rhat-target 0 $BARGS > gpurun_out/n_bench_$name.json 2> gpurun_out/n_bench_$name.err; }
BARGS="" run t2048_1000 FOCT_SLICE_TICKS=2048
BARGS="" run t8192_1000 FOCT_SLICE_TICKS=8192
BARGS="--profiles 1776" run t2048_1776 FOCT_SLICE_TICKS=2048
BARGS="--profiles 3552" run t4096_3552 A=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/n_bench_t*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("n_bench_")[1], "ms", round(d["ms_per_step"],1), "grad/s %.3e"%d["grad_per_s"], "frac %.3f"%d["roofline"]["frac"], "ess/s %.0f"%d["value"], d["roofline"]["launch"], "rhat_max %.3f"%d["quality"]["rhat_max"])
    except Exception as e:
        print(f, "failed", e, open(f.replace(".json",".err")).read()[-300:])
PY
timeout 900 python bench.py > gpurun_out/n_bench_default.json 2> gpurun_out/n_bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/n_bench_reference.json 2> gpurun_out/n_bench_reference.err; echo "ref rc=$?"
FOCT_FORCE_PAIR=1 python scripts/ncu_target.py 1000 > gpurun_out/n_target.txt 2>&1
FOCT_FORCE_PAIR=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:nuts2w_kernel -c 1 -o gpurun_out/n_ncu_nuts2w -f python scripts/ncu_target.py 1000 > gpurun_out/n_ncu.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/n_target.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/n_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/n_ncu_launch.log 2>&1
echo "launch list rc=$?"
python - <<'PY'
import json
for f in ("n_bench_default","n_bench_reference"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value","ms_per_step","grad_per_s","kernel_ms","quality","until_converged","e2e","e2e_with_draws","gpu_launches")}, (d.get("roofline") or {}).get("frac"), d.get("cpu_baseline"))
    except Exception as e:
        print(f, "failed", e, open(f"gpurun_out/{f}.err").read()[-600:])
PY
