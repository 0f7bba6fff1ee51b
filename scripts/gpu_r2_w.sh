#!/bin/bash
# round-2 GPU check W: per-Nn stack placement (Nn sweep), R shim with foct_R_summary, the three sampling kernels on the parity tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python scripts/config_sweep.py > gpurun_out/w_config_sweep.json 2> gpurun_out/w_config_sweep.err; echo "sweep rc=$?"
timeout 600 python -m pytest tests/test_gpu_rshim.py tests/test_gpu_parity.py tests/test_gpu_continue.py -m gpu -q --timeout 240 --timeout-method thread > gpurun_out/w_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/w_pytest.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/w_config_sweep.json").read())
for k,v in d.items(): print(k, {a: (round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ("wall_s","step_s","grad_per_s","tflops","min_ess_per_s","rhat_q99","launch","profiles")})
PY
