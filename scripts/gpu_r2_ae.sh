#!/bin/bash
# round-2 GPU check AE: latency kernel with one chain per CTA (a single profile on four SMs); parity and continuation tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python - <<'PY'
import time, os, numpy as np
from fitoct_b200 import _abi as abi, _lib as L, synth
S = synth.make_profiles(150, modulated_only=True)
cfg = abi.default_cfg(n_warmup=500, n_iter=1500, seed=1)
for Nn, n in ((10, 1), (10, 38), (10, 74), (10, 75), (10, 148)):
    kind = 0 if Nn > 0 else 1
    b = abi.make_problems_dense(S["x"], S["Y"][:n], S["UY"][:n], S["theta0"][:n], S["Sigma0"][:n], Nn=Nn, ids=S["ids"][:n])
    spec = abi.default_spec(kind)
    L.sample(kind, b, n, spec, cfg)
    ts = []
    for r in range(2):
        t = time.perf_counter(); o = L.sample(kind, b, n, spec, cfg); ts.append(time.perf_counter() - t)
    print("Nn", Nn, "profiles", n, "wall_s", ["%.4f" % t for t in ts], "leapfrogs", o["n_leapfrog"].sum(), "rhat_max %.4f" % np.nanmax(o["summary"][:, :max(Nn + 5, 3), 9]), flush=True)
PY
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_continue.py tests/test_gpu_rshim.py -m gpu -q -x --timeout 240 --timeout-method thread > gpurun_out/ae_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/ae_pytest.log
