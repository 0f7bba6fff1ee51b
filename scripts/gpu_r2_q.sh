#!/bin/bash
# round-2 GPU check Q: the build as shipped — parity suite, default bench line, DRAM traffic of the bench-size launch, launch list, configs sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/parity_metrics.jsonl
timeout 900 python -m pytest tests -m gpu -q --timeout 240 --timeout-method thread > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest.log
tail -4 gpurun_out/q_pytest.log
timeout 900 python bench.py > gpurun_out/q_bench_default.json 2> gpurun_out/q_bench_default.err; echo "bench rc=$?"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:nuts2w_kernel -c 1 --csv --log-file gpurun_out/q_dram_bench_launch.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --rhat-target 0 > gpurun_out/q_dram.log 2>&1; echo "dram rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/q_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/q_ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 600 python scripts/config_sweep.py > gpurun_out/q_config_sweep.json 2> gpurun_out/q_config_sweep.err; echo "sweep rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/q_bench_default.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value","ms_per_step","grad_per_s","kernel_ms","quality","until_converged","e2e","e2e_with_draws","gpu_launches")}, (d.get("roofline") or {}).get("frac"))
print(open("gpurun_out/q_config_sweep.json").read()[:3000])
print(open("gpurun_out/q_dram_bench_launch.csv").read()[-800:])
PY
