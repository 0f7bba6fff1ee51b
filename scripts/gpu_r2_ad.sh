#!/bin/bash
# round-2 GPU check AD: ncu --set full capture of the latency kernel (one profile x 4 chains, 60 + 40 iterations)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python scripts/ncu_target.py 1 > gpurun_out/ad_target.txt 2>&1; cat gpurun_out/ad_target.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nuts_lat_kernel -c 1 -o gpurun_out/ad_ncu_lat -f python scripts/ncu_target.py 1 > gpurun_out/ad_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/ad_ncu_lat.ncu-rep
