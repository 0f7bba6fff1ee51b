#!/bin/bash
# round-2 GPU check X: shared-basis mode of the one-chain-per-warp kernel (nuts_kernel<.., GB = 1>) against whole staged blobs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ls fitoct_b200/csrc/build 2>/dev/null | head -3
python - <<'PY'
import os, json, numpy as np
from fitoct_b200 import _abi as abi, _lib as L, synth
cfg = abi.default_cfg(n_warmup=500, n_iter=1500, seed=1)
for Nn, n in ((15, 888), (20, 592), (10, 444), (12, 444)):
    S = synth.make_profiles(n, modulated_only=True)
    b = abi.make_problems_dense(S["x"], S["Y"], S["UY"], S["theta0"], S["Sigma0"], Nn=Nn, ids=S["ids"])
    for mode in ("gb", "blob"):
        if mode == "blob": os.environ["FOCT_NO_SHARED_BASIS"] = "1"
        else: os.environ.pop("FOCT_NO_SHARED_BASIS", None)
        plan = L.Plan(0, b, n, abi.default_spec(), cfg, want_draws=False, want_summary=True)
        plan.run(1); plan.sync(); plan.run(2); plan.sync()
        tm = plan.timing(); o = plan.fetch(); plan.close()
        T = (tm["sample_ms"] + tm["summary_ms"]) * 1e-3
        leap = float(o["n_leapfrog"].sum())
        print("Nn", Nn, "profiles", n, mode, "step_s %.3f" % T, "grad/s %.4e" % (leap / T), "tflops %.2f" % (leap * 481 * (4 * Nn + 23) / T / 1e12),
              {k: tm[k] for k in ("grid", "block", "blocks_per_sm", "regs", "smem_bytes")}, "rhat_q99 %.4f" % np.nanquantile(np.nanmax(o["summary"][:, :Nn + 5, 9], axis=1), 0.99), flush=True)
PY
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_continue.py -m gpu -q -x --timeout 240 --timeout-method thread > gpurun_out/x_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/x_pytest.log
