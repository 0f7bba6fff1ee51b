#!/bin/bash
# round-2 GPU check G: kernel variants at full waves (unsliced) and at 1000 profiles (sliced); small batches pair vs one chain per warp
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
V=$PWD/fitoct_b200/variants
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --rhat-target 0 $BARGS > gpurun_out/g_bench_$name.json 2> gpurun_out/g_bench_$name.err; }
BARGS="--profiles 1776" run new_1776 FOCT_SLICE_TICKS=0
BARGS="--profiles 1776" run old_1776 FOCT_LIB_PATH=$V/lib_old.so
BARGS="--profiles 1776" run rl26_1776 FOCT_LIB_PATH=$V/lib_rl26.so FOCT_SLICE_TICKS=0
BARGS="--profiles 1776" run rl46_1776 FOCT_LIB_PATH=$V/lib_rl46.so FOCT_SLICE_TICKS=0
BARGS="--profiles 2368" run rl28_2368 FOCT_LIB_PATH=$V/lib_rl28.so FOCT_SLICE_TICKS=0
BARGS="--profiles 2368" run rl48_2368 FOCT_LIB_PATH=$V/lib_rl48.so FOCT_SLICE_TICKS=0
BARGS="" run rl28_1000 FOCT_LIB_PATH=$V/lib_rl28.so
BARGS="" run rl46_1000 FOCT_LIB_PATH=$V/lib_rl46.so
BARGS="" run rl48_1000 FOCT_LIB_PATH=$V/lib_rl48.so
BARGS="" run new_1000_s8192 FOCT_SLICE_TICKS=8192
BARGS="" run new_1000_s16384 FOCT_SLICE_TICKS=16384
for n in 1 16 148 444; do
  BARGS="--profiles $n" run pair_$n A=1
  BARGS="--profiles $n" run nopair_$n FOCT_NO_PAIR=1
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/g_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("g_bench_")[1], "ms", round(d["ms_per_step"],1), "grad/s %.3e"%d["grad_per_s"], "frac %.3f"%d["roofline"]["frac"], "ess/s %.0f"%d["value"], d["roofline"]["launch"], "rhat_max %.3f"%d["quality"]["rhat_max"])
    except Exception as e:
        print(f, "failed", e, open(f.replace(".json",".err")).read()[-300:])
PY
