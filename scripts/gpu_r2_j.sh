#!/bin/bash
# round-2 GPU check J (I hung on a 32-thread CTA at a 64-thread named barrier): basis copy in shared memory (sub-CTAs) against the L1 path, after the code-size cut
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/parity_metrics.jsonl
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/j_smoke.log
timeout 900 python -m pytest tests -m gpu -q -x --timeout 240 --timeout-method thread > gpurun_out/j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j_pytest.log
tail -8 gpurun_out/j_pytest.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --rhat-target 0 $BARGS > gpurun_out/j_bench_$name.json 2> gpurun_out/j_bench_$name.err; }
BARGS="--profiles 1776" run gs_1776 FOCT_SLICE_TICKS=0
BARGS="--profiles 1776" run l1_1776 FOCT_SLICE_TICKS=0 FOCT_BASIS_MODE=1
BARGS="--profiles 1776" run gs_1776_sliced A=1
BARGS="" run gs_1000 A=1
BARGS="" run l1_1000 FOCT_BASIS_MODE=1
BARGS="--profiles 888" run gs_888 A=1
BARGS="--profiles 888" run nopair_888 FOCT_NO_PAIR=1
BARGS="--profiles 750" run gs_750 FOCT_FORCE_PAIR=1
BARGS="--profiles 750" run nopair_750 FOCT_NO_PAIR=1
BARGS="--profiles 1776" run nin_1776 FOCT_SLICE_TICKS=0 FOCT_LIB_PATH=$PWD/fitoct_b200/variants/lib_nin.so
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/j_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("j_bench_")[1], "ms", round(d["ms_per_step"],1), "grad/s %.3e"%d["grad_per_s"], "frac %.3f"%d["roofline"]["frac"], "ess/s %.0f"%d["value"], d["roofline"]["launch"], "rhat_max %.3f"%d["quality"]["rhat_max"])
    except Exception as e:
        print(f, "failed", e, open(f.replace(".json",".err")).read()[-300:])
PY
