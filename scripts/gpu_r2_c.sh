#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
tail -15 gpurun_out/c_pytest.log
run() { # name, env..., args
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e $BARGS > gpurun_out/c_bench_$name.json 2> gpurun_out/c_bench_$name.err
}
BARGS="" run u4_1000 A=1
BARGS="--profiles 1184" run u4_1184 A=1
BARGS="" run u2_1000 FOCT_LIB_PATH=$PWD/fitoct_b200/variants/lib_u2.so
BARGS="--profiles 1184" run u2_1184 FOCT_LIB_PATH=$PWD/fitoct_b200/variants/lib_u2.so
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/c_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("c_bench_")[1], "ms", round(d["ms_per_step"],1), "grad/s %.3e"%d["grad_per_s"], "frac %.3f"%d["roofline"]["frac"], "ess/s %.0f"%d["value"], d["roofline"]["launch"]["regs"], "rhat_max %.3f"%d["quality"]["rhat_max"])
    except Exception as e:
        print(f, "failed", e)
PY
