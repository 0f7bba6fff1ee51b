#!/bin/bash
# round-2 GPU check R: software-pipelined basis loads (FOCT_PREFETCH), theta3 folded into the control values, exp table in
# shared memory, 8 CTAs/SM at 128 registers — A/B builds from scripts/build_variant.sh, 1000 and 1776 profiles each
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; lib=$2; shift 2; FOCT_LIB_PATH=$PWD/fitoct_b200/variants/lib_$lib.so timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --rhat-target 0 "$@" > gpurun_out/r_bench_$name.json 2> gpurun_out/r_bench_$name.err; }
for v in base0 pf pf_fold pf_et pf_all pf_b7; do
  run ${v}_1776 $v --profiles 1776
  run ${v}_1000 $v
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("r_bench_")[1], "ms", round(d["ms_per_step"],1), "grad/s %.4e"%d["grad_per_s"], "frac %.4f"%d["roofline"]["frac"], "ess/s %.0f"%d["value"], d["roofline"]["launch"], "rhat_max %.6f"%d["quality"]["rhat_max"], "ess %.4f" % d["quality"]["mean_min_bulk_ess_per_profile"])
    except Exception as e:
        print(f, "failed", e, open(f.replace(".json",".err")).read()[-300:])
PY
