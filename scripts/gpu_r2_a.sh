#!/bin/bash
# round-2 GPU check A: parity suite on the two-chains-per-warp kernel, then A/B bench against the one-chain kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/a_bench_pair.json 2> gpurun_out/a_bench_pair.err; echo "pair rc=$?"
FOCT_NO_PAIR=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/a_bench_nopair.json 2> gpurun_out/a_bench_nopair.err; echo "nopair rc=$?"
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --profiles 1184 > gpurun_out/a_bench_pair_1184.json 2>> gpurun_out/a_bench_pair.err
python - <<'PY'
import json
for f in ("a_bench_pair","a_bench_nopair","a_bench_pair_1184"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "ms", round(d["ms_per_step"],1), "grad/s %.3e"%d["grad_per_s"], "frac %.3f"%d["roofline"]["frac"], "ess/s %.0f"%d["value"], d["roofline"]["launch"], "rhat_max %.3f"%d["quality"]["rhat_max"])
    except Exception as e:
        print(f, "failed", e)
PY
