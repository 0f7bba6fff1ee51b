#!/bin/bash
# round-2 GPU check Z (2 GPUs): the multi-rank bench line of the shipped build incl. the in-library devices[] path and a small c5-shaped run
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/z_gpus.txt 2>&1
FOCT_TRACE=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 3 --c5 --c5-profiles 6000 > gpurun_out/z_bench_n2.json 2> gpurun_out/z_bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/z_bench_n2.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value","ms_per_step","n_gpus","grad_per_s","e2e","inlib","c5")})
PY
