"""ctypes mirror of include/fitoct_b200.h (structs and constants only; no library is loaded here).

Both the product binding (`fitoct_b200._lib`) and the test oracle wrapper (`oracle/oracle.py`) marshal
their arguments through these definitions so that the parity tests feed both sides identical bytes.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

FOCT_MAX_D = 32
FOCT_MAX_NN = 25
FOCT_MAX_CHAINS = 8
FOCT_N_SAMPLER_PARAMS = 6
FOCT_N_SUMMARY_COLS = 11
FOCT_EXPGP = 0
FOCT_MONOEXP = 1
FOCT_GRID_INTERNAL = 0
FOCT_GRID_EXTREMAL = 1

SAMPLER_PARAM_NAMES = ("accept_stat__", "stepsize__", "treedepth__", "n_leapfrog__", "divergent__", "energy__")
SUMMARY_COL_NAMES = ("mean", "se_mean", "sd", "2.5%", "25%", "50%", "75%", "97.5%", "n_eff", "Rhat", "Bulk_ESS")

c_double_p = C.POINTER(C.c_double)


class ModelSpec(C.Structure):
    _fields_ = [
        ("modulation", C.c_int),
        ("kernel", C.c_int),
        ("jitter", C.c_double),
        ("ygp_prior", C.c_int),
        ("lambda_prior", C.c_int),
        ("sigma_mean", C.c_double),
        ("sigma_sd", C.c_double),
        ("theta_prior", C.c_int),
        ("br_ndf", C.c_int),
    ]


class Problem(C.Structure):
    _fields_ = [
        ("N", C.c_int),
        ("x", c_double_p),
        ("y", c_double_p),
        ("uy", c_double_p),
        ("dataType", C.c_int),
        ("Nn", C.c_int),
        ("gridType", C.c_int),
        ("rho", C.c_double),
        ("lambda_rate", C.c_double),
        ("theta0", C.c_double * 3),
        ("Sigma0", C.c_double * 9),
        ("prior_PD", C.c_int),
        ("id", C.c_longlong),
    ]


class SamplerCfg(C.Structure):
    _fields_ = [
        ("chains", C.c_int),
        ("n_warmup", C.c_int),
        ("n_iter", C.c_int),
        ("adapt_delta", C.c_double),
        ("max_treedepth", C.c_int),
        ("stepsize0", C.c_double),
        ("seed", C.c_ulonglong),
        ("init_mode", C.c_int),
        ("init", c_double_p),
        ("save_warmup", C.c_int),
        ("gamma", C.c_double),
        ("kappa", C.c_double),
        ("t0", C.c_double),
        ("init_buffer", C.c_int),
        ("term_buffer", C.c_int),
        ("window", C.c_int),
        ("n_devices", C.c_int),
        ("devices", C.POINTER(C.c_int)),
        # ABI 2: continuation / run until converged
        ("inv_metric_init", c_double_p),
        ("stepsize_init", c_double_p),
        ("iter_offset", C.c_int),
        ("rhat_target", C.c_double),
        ("max_extend", C.c_int),
        ("extend_iter", C.c_int),
    ]


class Result(C.Structure):
    _fields_ = [
        ("draws", c_double_p),
        ("sampler_params", c_double_p),
        ("summary", c_double_p),
        ("stepsize", c_double_p),
        ("inv_metric", c_double_p),
        ("n_leapfrog", c_double_p),
        ("n_divergent", c_double_p),
        ("last_q", c_double_p),
        ("n_extend", C.POINTER(C.c_int)),
    ]


FOCT_PRIOR_MONO, FOCT_PRIOR_ABC = 0, 1
c_int_p = C.POINTER(C.c_int)


class PipelineCfg(C.Structure):
    """foct_pipeline_cfg: the ctrlParams.yaml keys the batch pipeline needs (FitOCT.R:37-53)."""
    _fields_ = [
        ("smooth_df", C.c_double),
        ("max_rate", C.c_double),
        ("prior_type", C.c_int),
        ("ru_theta", C.c_double),
        ("Nn", C.c_int),
        ("gridType", C.c_int),
        ("rho_scale", C.c_double),
        ("lambda_rate", C.c_double),
        ("gate", C.c_int),
    ]


class PipelineOut(C.Structure):
    _fields_ = [
        ("uy", c_double_p),
        ("ySmooth", c_double_p),
        ("noise_theta", c_double_p),
        ("mono_theta", c_double_p),
        ("mono_hessian", c_double_p),
        ("mono_br", c_double_p),
        ("mono_status", c_int_p),
        ("br_ci", c_double_p),
        ("alert", c_int_p),
        ("theta0", c_double_p),
        ("Sigma0", c_double_p),
        ("ru", c_double_p),
        ("n_expgp", C.c_int),
        ("expgp_index", c_int_p),
        ("expgp", Result),
        ("status", c_int_p),
    ]


class VbCfg(C.Structure):
    """foct_vb_cfg: rstan::vb's arguments (MODEL_SPEC §14)."""
    _fields_ = [
        ("iter", C.c_int),
        ("grad_samples", C.c_int),
        ("elbo_samples", C.c_int),
        ("eval_elbo", C.c_int),
        ("output_samples", C.c_int),
        ("adapt_engaged", C.c_int),
        ("adapt_iter", C.c_int),
        ("eta", C.c_double),
        ("tol_rel_obj", C.c_double),
        ("seed", C.c_ulonglong),
        ("init_mode", C.c_int),
        ("init", c_double_p),
        ("omega0", C.c_double),
    ]


class VbResult(C.Structure):
    _fields_ = [
        ("mean", c_double_p),
        ("draws", c_double_p),
        ("mu", c_double_p),
        ("omega", c_double_p),
        ("elbo", c_double_p),
        ("eta", c_double_p),
        ("iters", c_int_p),
        ("status", c_int_p),
    ]


def default_vb_cfg(**kw) -> VbCfg:
    """rstan::vb defaults; kept in step with foct_vb_cfg_default (tests/test_abi.py)."""
    c = VbCfg(iter=10000, grad_samples=1, elbo_samples=100, eval_elbo=100, output_samples=1000, adapt_engaged=1,
              adapt_iter=50, eta=1.0, tol_rel_obj=0.01, seed=1234, init_mode=0, omega0=0.0)
    for k, v in kw.items():
        if not hasattr(c, k):
            raise TypeError(f"unknown vb key {k}")
        setattr(c, k, v)
    return c


def alloc_vb_result(kind: int, n: int, Nn: int, cfg: VbCfg, draws=True):
    D, P_out = dims(kind, Nn)
    out = dict(mean=np.full((n, P_out), np.nan), draws=np.full((n, cfg.output_samples, P_out), np.nan) if draws else None,
               mu=np.full((n, D), np.nan), omega=np.full((n, D), np.nan), elbo=np.full(n, np.nan), eta=np.full(n, np.nan),
               iters=np.zeros(n, dtype=np.int32), status=np.full(n, -1, dtype=np.int32))
    R = VbResult()
    for k, v in out.items():
        if v is None:
            continue
        setattr(R, k, v.ctypes.data_as(c_int_p) if v.dtype == np.int32 else as_ptr(v))
    return out, R


def dims(kind: int, Nn: int) -> tuple[int, int]:
    """(D unconstrained dims, P_out output columns) — MODEL_SPEC §2, §6."""
    if kind == FOCT_EXPGP:
        return Nn + 5, Nn + 7
    return 3, 5


def param_names(kind: int, Nn: int) -> list[str]:
    """Output column names in the order plotExpGP.R:9,41 uses them."""
    if kind == FOCT_EXPGP:
        return (
            [f"theta[{i}]" for i in (1, 2, 3)]
            + [f"yGP[{k}]" for k in range(1, Nn + 1)]
            + ["lambda", "sigma", "br", "lp__"]
        )
    return [f"theta[{i}]" for i in (1, 2, 3)] + ["br", "lp__"]


def default_spec(kind: int = FOCT_EXPGP) -> ModelSpec:
    """MODEL_SPEC §8 defaults (same values as foct_model_spec_default)."""
    s = ModelSpec()
    s.modulation = 0
    s.kernel = 0
    s.jitter = 1e-9
    s.ygp_prior = 0
    s.lambda_prior = 0
    s.sigma_mean = 1.0
    s.sigma_sd = 0.1
    s.theta_prior = 0 if kind == FOCT_EXPGP else 1
    s.br_ndf = 0
    return s


def default_cfg(**kw) -> SamplerCfg:
    """rstan defaults (same values as foct_sampler_cfg_default); FitOCT.R:43-44 for warmup/iter."""
    c = SamplerCfg()
    c.chains = 4
    c.n_warmup = 500
    c.n_iter = 1500
    c.adapt_delta = 0.8
    c.max_treedepth = 10
    c.stepsize0 = 1.0
    c.seed = 1234
    c.init_mode = 0
    c.save_warmup = 0
    for k, v in kw.items():
        if not hasattr(c, k):
            raise TypeError(f"unknown sampler option {k!r}")
        setattr(c, k, v)
    return c


@dataclass
class ProblemBatch:
    """Owns the numpy buffers a ctypes `Problem[]` points into."""

    array: C.Array
    keep: list = field(default_factory=list)

    def __len__(self) -> int:
        return len(self.array)


def make_problems(profiles: list[dict]) -> ProblemBatch:
    """profiles: dicts with x, y, uy, dataType, Nn, gridType, rho, lambda_rate, theta0, Sigma0, prior_PD, id."""
    arr = (Problem * max(len(profiles), 1))()
    keep = []
    for j, p in enumerate(profiles):
        x = np.ascontiguousarray(p["x"], dtype=np.float64)
        y = np.ascontiguousarray(p["y"], dtype=np.float64)
        uy = np.ascontiguousarray(p["uy"], dtype=np.float64)
        if not (x.shape == y.shape == uy.shape and x.ndim == 1):
            raise ValueError("x, y, uy must be 1-D arrays of equal length")
        keep += [x, y, uy]
        q = arr[j]
        q.N = x.shape[0]
        q.x = x.ctypes.data_as(c_double_p)
        q.y = y.ctypes.data_as(c_double_p)
        q.uy = uy.ctypes.data_as(c_double_p)
        q.dataType = int(p.get("dataType", 2))
        q.Nn = int(p.get("Nn", 10))
        g = p.get("gridType", "internal")
        q.gridType = {"internal": 0, "extremal": 1}.get(g, g) if isinstance(g, str) else int(g)
        q.rho = float(p.get("rho", 1.0 / max(q.Nn, 1)))
        q.lambda_rate = float(p.get("lambda_rate", 0.1))
        th0 = np.asarray(p.get("theta0", (0.0, 0.0, 1.0)), dtype=np.float64).reshape(3)
        S0 = np.asarray(p.get("Sigma0", np.eye(3)), dtype=np.float64).reshape(9)
        for i in range(3):
            q.theta0[i] = th0[i]
        for i in range(9):
            q.Sigma0[i] = S0[i]
        q.prior_PD = int(p.get("prior_PD", 0))
        q.id = int(p.get("id", j))
    return ProblemBatch(arr, keep)


def as_ptr(a: np.ndarray | None):
    if a is None:
        return c_double_p()
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_double_p)


# numpy view of `Problem[]` for filling large batches without a Python loop
PROBLEM_DTYPE = np.dtype(
    [
        ("N", np.int32),
        ("x", np.uintp),
        ("y", np.uintp),
        ("uy", np.uintp),
        ("dataType", np.int32),
        ("Nn", np.int32),
        ("gridType", np.int32),
        ("rho", np.float64),
        ("lambda_rate", np.float64),
        ("theta0", np.float64, (3,)),
        ("Sigma0", np.float64, (9,)),
        ("prior_PD", np.int32),
        ("id", np.int64),
    ],
    align=True,
)
assert PROBLEM_DTYPE.itemsize == C.sizeof(Problem), (PROBLEM_DTYPE.itemsize, C.sizeof(Problem))


def make_problems_dense(x, Y, UY, theta0, Sigma0, *, dataType=2, Nn=10, gridType=0, rho=None, lambda_rate=0.1,
                        prior_PD=0, ids=None) -> ProblemBatch:
    """Batch of equal-length profiles. x: [N] shared or [n,N]; Y, UY: [n,N]; theta0: [n,3]; Sigma0: [n,3,3]."""
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    UY = np.ascontiguousarray(UY, dtype=np.float64)
    n, N = Y.shape
    x = np.ascontiguousarray(x, dtype=np.float64)
    X = np.ascontiguousarray(np.broadcast_to(x, (n, N))) if x.ndim == 1 else x
    arr = (Problem * max(n, 1))()
    v = np.frombuffer(arr, dtype=PROBLEM_DTYPE, count=max(n, 1))
    if n:
        v = v[:n]
        v["N"] = N
        v["x"] = X.ctypes.data + np.arange(n, dtype=np.uintp) * np.uintp(N * 8)
        v["y"] = Y.ctypes.data + np.arange(n, dtype=np.uintp) * np.uintp(N * 8)
        v["uy"] = UY.ctypes.data + np.arange(n, dtype=np.uintp) * np.uintp(N * 8)
        v["dataType"] = dataType
        v["Nn"] = Nn
        v["gridType"] = gridType
        v["rho"] = (1.0 / max(Nn, 1)) if rho is None else rho
        v["lambda_rate"] = lambda_rate
        v["theta0"] = np.asarray(theta0, dtype=np.float64).reshape(n, 3)
        v["Sigma0"] = np.asarray(Sigma0, dtype=np.float64).reshape(n, 9)
        v["prior_PD"] = prior_PD
        v["id"] = np.arange(n, dtype=np.int64) if ids is None else np.asarray(ids, dtype=np.int64)
    return ProblemBatch(arr, [X, Y, UY])
