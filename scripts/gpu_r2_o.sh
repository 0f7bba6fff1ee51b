#!/bin/bash
# round-2 GPU check O (2 GPUs): the multi-rank bench line incl. the in-library devices[] path and a small c5-shaped run; 2-GPU tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/o_gpus.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 3 --c5 --c5-profiles 6000 > gpurun_out/o_bench_n2.json 2> gpurun_out/o_bench_n2.err; echo "bench n2 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/o_bench_ref_n2.json 2> gpurun_out/o_bench_ref_n2.err; echo "ref n2 rc=$?"
timeout 600 python -m pytest tests -m gpu -q --timeout 240 --timeout-method thread -k "shard or plan_path or multi" > gpurun_out/o_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/o_pytest.log
python - <<'PY'
import json
for f in ("o_bench_n2","o_bench_ref_n2"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ("value","n_gpus","ms_per_step","grad_per_s","quality","until_converged","e2e","inlib","c5","gpu_launches")}, (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "failed", e, open(f"gpurun_out/{f}.err").read()[-1500:])
PY
