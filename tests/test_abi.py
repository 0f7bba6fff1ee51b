"""CPU: the C-ABI library loads and exports every symbol include/fitoct_b200.h declares; struct layouts of the
ctypes mirror match the header; without a GPU every compute entry point fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from fitoct_b200 import _abi as abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fitoct_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(foct_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported():
    from fitoct_b200 import _lib

    L = _lib.lib()
    names = header_functions()
    assert len(names) >= 20
    assert set(names) == set(_lib.EXPORTS)
    for n in names:
        assert hasattr(L, n), n
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for n in names:
        assert re.search(rf"\bT {n}\b", nm), f"{n} not exported with C linkage"


def test_struct_layouts_match_header():
    # compile a tiny C program against the public header and compare sizeof/offsetof with the ctypes mirror
    import tempfile

    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "fitoct_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu\n", sizeof(foct_model_spec), sizeof(foct_problem), sizeof(foct_sampler_cfg), sizeof(foct_result));
  printf("%zu %zu %zu %zu\n", offsetof(foct_problem, theta0), offsetof(foct_problem, id), offsetof(foct_sampler_cfg, seed), offsetof(foct_sampler_cfg, devices));
  printf("%d %d %d %d\n", FOCT_MAX_NN, FOCT_MAX_CHAINS, FOCT_N_SUMMARY_COLS, FOCT_N_SAMPLER_PARAMS);
  printf("%zu %zu %zu %zu %zu\n", sizeof(foct_pipeline_cfg), sizeof(foct_pipeline_out), offsetof(foct_pipeline_cfg, gate),
         offsetof(foct_pipeline_out, n_expgp), offsetof(foct_pipeline_out, expgp));
  printf("%zu %zu %zu %zu\n", sizeof(foct_vb_cfg), sizeof(foct_vb_result), offsetof(foct_vb_cfg, seed), offsetof(foct_vb_cfg, omega0));
  return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write(prog)
        exe = os.path.join(d, "t")
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.run([cc, "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    vals = [int(v) for v in out]
    assert vals[:4] == [C.sizeof(abi.ModelSpec), C.sizeof(abi.Problem), C.sizeof(abi.SamplerCfg), C.sizeof(abi.Result)]
    assert vals[4:8] == [abi.Problem.theta0.offset, abi.Problem.id.offset, abi.SamplerCfg.seed.offset,
                         abi.SamplerCfg.devices.offset]
    assert vals[8:12] == [abi.FOCT_MAX_NN, abi.FOCT_MAX_CHAINS, abi.FOCT_N_SUMMARY_COLS, abi.FOCT_N_SAMPLER_PARAMS]
    assert vals[12:17] == [C.sizeof(abi.PipelineCfg), C.sizeof(abi.PipelineOut), abi.PipelineCfg.gate.offset,
                           abi.PipelineOut.n_expgp.offset, abi.PipelineOut.expgp.offset]
    assert vals[17:] == [C.sizeof(abi.VbCfg), C.sizeof(abi.VbResult), abi.VbCfg.seed.offset, abi.VbCfg.omega0.offset]


def test_defaults_and_dims_without_device():
    from fitoct_b200 import _lib

    L = _lib.lib()
    v, w = abi.VbCfg(), abi.default_vb_cfg()
    L.foct_vb_cfg_default(C.byref(v))
    for name, _ in abi.VbCfg._fields_:
        if name != "init":
            assert getattr(v, name) == getattr(w, name), name
    assert L.foct_version() == 2
    for kind in (abi.FOCT_EXPGP, abi.FOCT_MONOEXP):
        s = abi.ModelSpec()
        L.foct_model_spec_default(C.byref(s), kind)
        ref = abi.default_spec(kind)
        for f, _ in abi.ModelSpec._fields_:
            assert getattr(s, f) == getattr(ref, f), f
    c = abi.SamplerCfg()
    L.foct_sampler_cfg_default(C.byref(c))
    ref = abi.default_cfg()
    for f in ("chains", "n_warmup", "n_iter", "adapt_delta", "max_treedepth", "stepsize0", "seed", "init_mode", "save_warmup"):
        assert getattr(c, f) == getattr(ref, f), f
    D, P = C.c_int(), C.c_int()
    assert L.foct_dims(abi.FOCT_EXPGP, 10, C.byref(D), C.byref(P)) == 0 and (D.value, P.value) == (15, 17)
    assert L.foct_dims(abi.FOCT_MONOEXP, 0, C.byref(D), C.byref(P)) == 0 and (D.value, P.value) == (3, 5)
    assert L.foct_dims(abi.FOCT_EXPGP, 26, C.byref(D), C.byref(P)) != 0
    assert b"Nn=26" in L.foct_last_error()
    np.testing.assert_allclose(_lib.grid(10, 0), np.linspace(1 / 22, 1 - 1 / 22, 10), atol=1e-15)


def test_no_cpu_fallback():
    from fitoct_b200 import _lib

    if _lib.device_count() > 0:
        pytest.skip("a GPU is present; the no-device error path is exercised on CPU-only boxes")
    from fitoct_b200 import synth

    S = synth.make_profiles(1)
    b = abi.make_problems_dense(S["x"], S["Y"], S["UY"], S["theta0"], S["Sigma0"], Nn=5)
    with pytest.raises(_lib.FitOCTError) as ei:
        _lib.sample(abi.FOCT_EXPGP, b, 1, abi.default_spec(), abi.default_cfg(n_warmup=5, n_iter=10))
    assert ei.value.code == -2 and "no CPU path" in str(ei.value)
    with pytest.raises(_lib.FitOCTError):
        _lib.logp_grad(abi.FOCT_EXPGP, b, 1, abi.default_spec(), np.zeros((1, 1, 10)))
    with pytest.raises(_lib.FitOCTError):
        _lib.monoexp_map(b, 1, abi.default_spec(abi.FOCT_MONOEXP))


def test_product_never_imports_oracle():
    # the product package must not reference oracle/ anywhere (a CPU fallback would void every parity claim)
    pkg = os.path.join(ROOT, "fitoct_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "foct_oracle" not in txt, f
    # nor may the helper scripts: only tests/, smoke() and bench.py's baseline legs may execute the oracle
    for f in os.listdir(os.path.join(ROOT, "scripts")):
        if f.endswith(".py"):
            txt = open(os.path.join(ROOT, "scripts", f)).read()
            assert "import oracle" not in txt and "from oracle" not in txt, f
