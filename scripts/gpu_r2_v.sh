#!/bin/bash
# round-2 GPU check V: the build as shipped (pipelined basis loads, exp table and stack level 0 in shared memory, c x shared, latency kernel) — parity suite,
# default bench line, reference arm, DRAM traffic of the bench-size launch, ncu --set full capture, launch list, configs sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/parity_metrics.jsonl
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/v_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/v_smoke.log
timeout 900 python -m pytest tests -m gpu -q --timeout 240 --timeout-method thread > gpurun_out/v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/v_pytest.log
tail -4 gpurun_out/v_pytest.log
timeout 900 python bench.py > gpurun_out/v_bench_default.json 2> gpurun_out/v_bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/v_bench_reference.json 2> gpurun_out/v_bench_reference.err; echo "ref rc=$?"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:nuts2w_kernel -c 1 --csv --log-file gpurun_out/v_dram_bench_launch.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --rhat-target 0 > gpurun_out/v_dram.log 2>&1; echo "dram rc=$?"
FOCT_FORCE_PAIR=1 python scripts/ncu_target.py 1000 > gpurun_out/v_target.txt 2>&1
FOCT_FORCE_PAIR=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:nuts2w_kernel -c 1 -o gpurun_out/v_ncu_nuts2w -f python scripts/ncu_target.py 1000 > gpurun_out/v_ncu.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/v_target.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/v_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/v_ncu_launch.log 2>&1
echo "launch list rc=$?"
timeout 600 python scripts/config_sweep.py > gpurun_out/v_config_sweep.json 2> gpurun_out/v_config_sweep.err; echo "sweep rc=$?"
python - <<'PY'
import json
for f in ("v_bench_default","v_bench_reference"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value","ms_per_step","grad_per_s","kernel_ms","quality","until_converged","e2e","e2e_with_draws","gpu_launches")}, (d.get("roofline") or {}).get("frac"), d.get("cpu_baseline"))
    except Exception as e:
        print(f, "failed", e, open(f"gpurun_out/{f}.err").read()[-600:])
print(open("gpurun_out/v_config_sweep.json").read()[:3500])
print(open("gpurun_out/v_config_sweep.err").read()[-500:])
PY
