#!/bin/bash
# round-2 GPU check AA: full GPU test suite of the final tree; slice length and kernel threshold probes with the final kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/parity_metrics.jsonl
timeout 900 python -m pytest tests -m gpu -q --timeout 240 --timeout-method thread > gpurun_out/aa_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/aa_pytest.log
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --rhat-target 0 $BARGS > gpurun_out/aa_bench_$name.json 2> gpurun_out/aa_bench_$name.err; }
BARGS="" run t2048_1000 FOCT_SLICE_TICKS=2048
BARGS="" run t8192_1000 FOCT_SLICE_TICKS=8192
BARGS="" run t4096_1000 A=1
BARGS="--profiles 600" run p600_auto A=1
BARGS="--profiles 600" run p600_pair FOCT_FORCE_PAIR=1
BARGS="--profiles 720" run p720_auto A=1
BARGS="--profiles 720" run p720_one FOCT_NO_PAIR=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/aa_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("aa_bench_")[1], "ms", round(d["ms_per_step"],1), "grad/s %.4e"%d["grad_per_s"], "frac %.4f"%d["roofline"]["frac"], d["roofline"]["launch"])
    except Exception as e:
        print(f, "failed", e, open(f.replace(".json",".err")).read()[-300:])
PY
