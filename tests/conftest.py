import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (test infrastructure)."""
    from oracle import oracle

    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "model_cases.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def L():
    """The CUDA product library; GPU tests fail loudly (not skip) if it or the device is missing."""
    from fitoct_b200 import _lib

    _lib.lib()
    assert _lib.device_count() >= 1, "GPU tests need a CUDA device: fitoct_b200 has no CPU path"
    return _lib


@pytest.fixture(params=["auto", "pair", "one"])
def sampling_kernel(request, monkeypatch):
    """The library picks the sampling kernel by batch size (foct_inst.cu): the latency kernel (two warps per chain) while
    there is at most one work item per SM, one chain per warp while the batch fits the GPU in one go, two chains per warp
    above that.  The parity tests use small batches, so they run three times: as shipped (= the latency kernel), with the
    two-chains-per-warp kernel forced — the one the BASELINE-size batches run on — and with the one-chain-per-warp kernel."""
    monkeypatch.delenv("FOCT_FORCE_PAIR", raising=False)
    monkeypatch.delenv("FOCT_NO_LAT", raising=False)
    if request.param == "pair":
        monkeypatch.setenv("FOCT_FORCE_PAIR", "1")
    elif request.param == "one":
        monkeypatch.setenv("FOCT_NO_LAT", "1")
    return request.param


def case_to_batch(case):
    from fitoct_b200 import _abi as abi

    spec = abi.ModelSpec()
    for k, v in case["spec"].items():
        setattr(spec, k, v)
    prof = dict(x=case["x"], y=case["y"], uy=case["uy"], dataType=case["dataType"], Nn=case["Nn"],
                gridType=case["gridType"], rho=case["rho"], lambda_rate=case["lambda_rate"], theta0=case["theta0"],
                Sigma0=case["Sigma0"], prior_PD=case["prior_PD"], id=0)
    return abi.make_problems([prof]), spec


def grad_tol_ok(g, g_ref, abs_terms, rtol):
    """|g - g_ref| <= rtol * (|g_ref| + sum_i |summand_i|): relative to the conditioning of each sum."""
    return np.all(np.abs(np.asarray(g) - np.asarray(g_ref)) <= rtol * (np.abs(g_ref) + abs_terms))


def record_metric(name, **values):
    """Append measured parity figures (worst errors, not just pass/fail) to gpurun_out/parity_metrics.jsonl so that the
    numbers behind the tolerances can be copied into profiles/ after a GPU run."""
    d = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_metrics.jsonl"), "a") as fh:
            fh.write(json.dumps(dict(test=name, **{k: (float(v) if isinstance(v, (int, float, np.integer, np.floating)) else v) for k, v in values.items()})) + "\n")
    except OSError:
        pass
