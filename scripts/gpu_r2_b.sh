#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -25 gpurun_out/b_pytest.log
python scripts/ncu_target.py 888 > gpurun_out/b_target_pair.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nuts2_kernel -c 1 -o gpurun_out/b_ncu_pair -f python scripts/ncu_target.py 888 > gpurun_out/b_ncu_pair.log 2>&1
echo "ncu pair rc=$?"; cat gpurun_out/b_target_pair.txt
FOCT_NO_PAIR=1 python scripts/ncu_target.py 444 > gpurun_out/b_target_nopair.txt 2>&1
FOCT_NO_PAIR=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:nuts_kernel -c 1 -o gpurun_out/b_ncu_nopair -f python scripts/ncu_target.py 444 > gpurun_out/b_ncu_nopair.log 2>&1
echo "ncu nopair rc=$?"; cat gpurun_out/b_target_nopair.txt
ls -la gpurun_out/*.ncu-rep
