#!/bin/bash
# round-2 GPU check M: progress-aware yielding, L1 capacity hypothesis
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/parity_metrics.jsonl
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/m_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/m_smoke.log
timeout 900 python -m pytest tests -m gpu -q --timeout 240 --timeout-method thread > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/m_pytest.log
tail -8 gpurun_out/m_pytest.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --rhat-target 0 $BARGS > gpurun_out/m_bench_$name.json 2> gpurun_out/m_bench_$name.err; }
BARGS="" run prio_1000_a A=1
BARGS="--seed 99" run prio_1000_b A=1
BARGS="--seed 7" run prio_1000_c A=1
BARGS="" run rr_1000_a FOCT_SLICE_PRIO=0
BARGS="--seed 7" run rr_1000_c FOCT_SLICE_PRIO=0
BARGS="--profiles 1776" run prio_1776 A=1
BARGS="--profiles 1776" run cta_1776_sliced FOCT_CTA_ITEMS=1
BARGS="--profiles 1776" run cta_1776_pad FOCT_CTA_ITEMS=1 FOCT_SMEM_PAD=12288
BARGS="--profiles 1332" run prio_1332 A=1
BARGS="--profiles 1332" run cta_1332 FOCT_CTA_ITEMS=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/m_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("m_bench_")[1], "ms", round(d["ms_per_step"],1), "grad/s %.3e"%d["grad_per_s"], "frac %.3f"%d["roofline"]["frac"], "ess/s %.0f"%d["value"], d["roofline"]["launch"], "rhat_max %.3f"%d["quality"]["rhat_max"])
    except Exception as e:
        print(f, "failed", e, open(f.replace(".json",".err")).read()[-300:])
PY
