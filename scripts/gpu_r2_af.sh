#!/bin/bash
# round-2 GPU check AF: final bench line, launch list and configs sweep of the shipped build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 python bench.py > gpurun_out/af_bench_default.json 2> gpurun_out/af_bench_default.err; echo "bench rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/af_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/af_ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 400 python scripts/config_sweep.py > gpurun_out/af_config_sweep.json 2> gpurun_out/af_config_sweep.err; echo "sweep rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/af_bench_default.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value","ms_per_step","grad_per_s","until_converged","e2e","single_profile")}, d["roofline"]["frac"])
s=json.loads(open("gpurun_out/af_config_sweep.json").read())
for k,v in s.items(): print(k, {a: (round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ("wall_s","step_s","grad_per_s","profiles")})
PY
