#!/bin/bash
# round-2 GPU check P: packed scalar slot of the subtree stack
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/p_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/p_smoke.log
timeout 600 python -m pytest tests/test_gpu_continue.py tests/test_gpu_parity.py -m gpu -q --timeout 240 --timeout-method thread > gpurun_out/p_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/p_pytest.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --rhat-target 0 $BARGS > gpurun_out/p_bench_$name.json 2> gpurun_out/p_bench_$name.err; }
BARGS="" run sc_1000_a A=1
BARGS="--seed 99" run sc_1000_b A=1
BARGS="--profiles 1776" run sc_1776 A=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/p_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("p_bench_")[1], "ms", round(d["ms_per_step"],1), "grad/s %.3e"%d["grad_per_s"], "frac %.3f"%d["roofline"]["frac"], "ess/s %.0f"%d["value"], d["roofline"]["launch"], "rhat_max %.3f"%d["quality"]["rhat_max"])
    except Exception as e:
        print(f, "failed", e, open(f.replace(".json",".err")).read()[-300:])
PY
