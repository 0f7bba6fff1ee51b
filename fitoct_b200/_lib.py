"""ctypes binding of libfitoct_b200.so (the CUDA product library behind include/fitoct_b200.h).

The library must be built in-tree (`make -C fitoct_b200/csrc`, or `__graft_entry__.build()`); there is no
CPU fallback — a missing library or a missing CUDA device is a hard error.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _abi as abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# FOCT_LIB_PATH: load another build of the same library (kernel A/B runs: scripts/build_variant.sh)
LIB_PATH = os.environ.get("FOCT_LIB_PATH") or os.path.join(_HERE, "libfitoct_b200.so")

# every symbol include/fitoct_b200.h declares (tests/test_abi.py checks the header against this list)
EXPORTS = (
    "foct_version", "foct_device_count", "foct_last_error", "foct_model_spec_default", "foct_sampler_cfg_default",
    "foct_dims", "foct_expgp_grid", "foct_expgp_basis", "foct_logp_grad", "foct_sample", "foct_expgp_sample",
    "foct_monoexp_sample", "foct_monoexp_map", "foct_expgp_map", "foct_predict", "foct_plan_create", "foct_plan_run",
    "foct_plan_sync", "foct_plan_timing", "foct_plan_fetch", "foct_plan_destroy", "foct_fp64_peak",
    "foct_estimate_noise", "foct_birge_ci", "foct_print_br", "foct_estimate_exp_prior", "foct_pipeline_cfg_default",
    "foct_pipeline", "foct_vb_cfg_default", "foct_vb", "foct_release_cache",
    "foct_sample_cb", "foct_plan_query", "foct_plan_cancel", "foct_expgp_logp_grad", "foct_plan_launches", "foct_summary",
)

PROGRESS_FN = C.CFUNCTYPE(C.c_int, C.c_double, C.c_char_p, C.c_void_p)

_LIB = None


class FitOCTError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"fitoct_b200 error {code}: {msg}")
        self.code = code


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build the CUDA extension with `make -C fitoct_b200/csrc -j` "
                "(fitoct_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        dp, ip = abi.c_double_p, C.POINTER(C.c_int)
        PP, MS, SC, RS = (C.POINTER(abi.Problem), C.POINTER(abi.ModelSpec), C.POINTER(abi.SamplerCfg),
                          C.POINTER(abi.Result))
        L.foct_last_error.restype = C.c_char_p
        L.foct_model_spec_default.argtypes = [MS, C.c_int]
        L.foct_model_spec_default.restype = None
        L.foct_sampler_cfg_default.argtypes = [SC]
        L.foct_sampler_cfg_default.restype = None
        L.foct_dims.argtypes = [C.c_int, C.c_int, ip, ip]
        L.foct_expgp_grid.argtypes = [C.c_int, C.c_int, dp]
        L.foct_expgp_basis.argtypes = [PP, MS, dp]
        L.foct_logp_grad.argtypes = [C.c_int, PP, C.c_int, MS, dp, C.c_int, dp, dp, dp]
        for name in ("foct_sample",):
            getattr(L, name).argtypes = [C.c_int, PP, C.c_int, MS, SC, RS]
        L.foct_sample_cb.argtypes = [C.c_int, PP, C.c_int, MS, SC, RS, PROGRESS_FN, C.c_void_p, C.c_int]
        L.foct_expgp_logp_grad.argtypes = [PP, C.c_int, MS, dp, C.c_int, dp, dp, dp]
        L.foct_plan_query.argtypes = [C.c_void_p, ip, dp]
        L.foct_plan_cancel.argtypes = [C.c_void_p]
        L.foct_plan_launches.argtypes = [C.c_void_p]
        for name in ("foct_expgp_sample", "foct_monoexp_sample"):
            getattr(L, name).argtypes = [PP, C.c_int, MS, SC, RS]
        L.foct_monoexp_map.argtypes = [PP, C.c_int, MS, dp, dp, dp, dp, ip]
        L.foct_expgp_map.argtypes = [PP, C.c_int, MS, dp, dp, dp, ip]
        L.foct_predict.argtypes = [C.c_int, PP, MS, dp, C.c_int, dp, dp, dp]
        L.foct_summary.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_int, dp]
        L.foct_plan_create.argtypes = [C.c_int, PP, C.c_int, MS, SC, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.foct_plan_run.argtypes = [C.c_void_p, C.c_ulonglong]
        L.foct_plan_sync.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.foct_plan_fetch.argtypes = [C.c_void_p, RS]
        L.foct_plan_timing.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)] + [ip] * 5
        L.foct_plan_destroy.argtypes = [C.c_void_p]
        L.foct_plan_destroy.restype = None
        L.foct_fp64_peak.argtypes = [C.c_int, dp, dp]
        L.foct_estimate_noise.argtypes = [PP, C.c_int, C.c_double, C.c_double, dp, dp, dp, dp, ip]
        L.foct_birge_ci.argtypes = [C.c_double, dp]
        L.foct_print_br.argtypes = [C.c_int, PP, C.c_int, MS, dp, dp, ip]
        L.foct_estimate_exp_prior.argtypes = [PP, C.c_int, C.c_int, dp, dp, C.c_double, dp, dp, dp]
        L.foct_pipeline_cfg_default.argtypes = [C.POINTER(abi.PipelineCfg)]
        L.foct_pipeline_cfg_default.restype = None
        L.foct_release_cache.argtypes = []
        L.foct_release_cache.restype = None
        L.foct_vb_cfg_default.argtypes = [C.POINTER(abi.VbCfg)]
        L.foct_vb_cfg_default.restype = None
        L.foct_vb.argtypes = [C.c_int, PP, C.c_int, MS, C.POINTER(abi.VbCfg), C.POINTER(abi.VbResult)]
        L.foct_pipeline.argtypes = [PP, C.c_int, C.POINTER(abi.PipelineCfg), MS, SC, C.POINTER(abi.PipelineOut)]
        _LIB = L
    return _LIB


def check(rc: int):
    if rc != 0:
        raise FitOCTError(rc, lib().foct_last_error().decode())


def device_count() -> int:
    return lib().foct_device_count()


def grid(Nn: int, gridType: int) -> np.ndarray:
    out = np.empty(Nn)
    check(lib().foct_expgp_grid(Nn, gridType, abi.as_ptr(out)))
    return out


def basis(batch: abi.ProblemBatch, j: int, spec: abi.ModelSpec) -> np.ndarray:
    p = batch.array[j]
    B = np.empty((p.Nn, p.N))
    check(lib().foct_expgp_basis(C.byref(p), C.byref(spec), abi.as_ptr(B)))
    return B


def logp_grad(kind: int, batch: abi.ProblemBatch, n_problems: int, spec: abi.ModelSpec, q: np.ndarray):
    """q: [n_problems, n_q, D] -> lp [n_problems, n_q], grad [n_problems, n_q, D], chi2 [n_problems, n_q]."""
    q = np.ascontiguousarray(q, dtype=np.float64)
    n, n_q, D = q.shape
    assert n == n_problems
    lp = np.empty((n, n_q))
    g = np.empty((n, n_q, D))
    chi2 = np.empty((n, n_q))
    check(lib().foct_logp_grad(kind, batch.array, n, C.byref(spec), abi.as_ptr(q), n_q, abi.as_ptr(lp), abi.as_ptr(g),
                               abi.as_ptr(chi2)))
    return lp, g, chi2


def alloc_result(kind, n_problems, Nn, cfg: abi.SamplerCfg, draws=True, summary=True):
    D, P_out = abi.dims(kind, Nn)
    n_saved = max(0, cfg.n_iter if cfg.save_warmup else cfg.n_iter - cfg.n_warmup)  # the library validates the cfg
    Cn = max(0, cfg.chains)
    out = dict(
        draws=np.full((n_problems, n_saved, Cn, P_out), np.nan) if draws else None,
        sampler_params=np.full((n_problems, n_saved, Cn, 6), np.nan) if draws else None,
        summary=np.full((n_problems, P_out, abi.FOCT_N_SUMMARY_COLS), np.nan) if summary else None,
        stepsize=np.full((n_problems, Cn), np.nan),
        inv_metric=np.full((n_problems, Cn, D), np.nan),
        n_leapfrog=np.zeros((n_problems, Cn, 2)),
        n_divergent=np.zeros((n_problems, Cn)),
        last_q=np.full((n_problems, Cn, D), np.nan),
        n_extend=np.zeros(n_problems, dtype=np.int32),
    )
    R = abi.Result()
    for k, v in out.items():
        setattr(R, k, v.ctypes.data_as(C.POINTER(C.c_int)) if v is not None and v.dtype == np.int32 else abi.as_ptr(v))
    return out, R


def continuation_cfg(cfg: abi.SamplerCfg, prev: dict, n_more: int, iters_done: int, keep: list) -> abi.SamplerCfg:
    """Sampler options that continue the run whose result is `prev` for `n_more` draws: no warm-up, the adapted metric
    and step size, start = the last state, Philox sites after the `iters_done` iterations already made.  `keep`
    receives the arrays the returned struct points into."""
    c = abi.SamplerCfg()
    C.memmove(C.byref(c), C.byref(cfg), C.sizeof(abi.SamplerCfg))
    init = np.ascontiguousarray(prev["last_q"], dtype=np.float64)
    invm = np.ascontiguousarray(prev["inv_metric"], dtype=np.float64)
    eps = np.ascontiguousarray(prev["stepsize"], dtype=np.float64)
    keep += [init, invm, eps]
    c.n_warmup, c.n_iter, c.save_warmup = 0, int(n_more), 0
    c.init_mode, c.init = 2, abi.as_ptr(init)
    c.inv_metric_init, c.stepsize_init = abi.as_ptr(invm), abi.as_ptr(eps)
    c.iter_offset = int(iters_done)
    c.rhat_target, c.max_extend, c.extend_iter = 0.0, 0, 0
    return c


def sample(kind: int, batch: abi.ProblemBatch, n_problems: int, spec: abi.ModelSpec, cfg: abi.SamplerCfg, draws=True,
           summary=True, devices=None, progress=None, poll_ms=100):
    """One-shot batched NUTS through the C ABI with host buffers (the `e2e` path).  progress(fraction, phase) -> truthy to
    cancel: called on this thread while the kernels run (foct_sample_cb); a cancelled run raises FitOCTError(-5)."""
    Nn = batch.array[0].Nn if kind == abi.FOCT_EXPGP else 0
    out, R = alloc_result(kind, n_problems, Nn, cfg, draws, summary)
    keep = None
    if devices is not None:
        keep = (C.c_int * len(devices))(*devices)
        cfg.n_devices = len(devices)
        cfg.devices = C.cast(keep, C.POINTER(C.c_int))
    try:
        if progress is None:
            check(lib().foct_sample(kind, batch.array, n_problems, C.byref(spec), C.byref(cfg), C.byref(R)))
        else:
            cb = PROGRESS_FN(lambda f, phase, _u: 1 if progress(float(f), phase.decode()) else 0)
            check(lib().foct_sample_cb(kind, batch.array, n_problems, C.byref(spec), C.byref(cfg), C.byref(R), cb, None, int(poll_ms)))
    finally:
        if devices is not None:
            cfg.n_devices = 0
            cfg.devices = C.POINTER(C.c_int)()
    return out


class Plan:
    """Device-resident batch (foct_plan_*): upload once, run/time the sampling kernel repeatedly."""

    def __init__(self, kind, batch, n_problems, spec, cfg, want_draws=False, want_summary=True):
        self.kind, self.n, self.cfg = kind, n_problems, cfg
        self.Nn = batch.array[0].Nn if kind == abi.FOCT_EXPGP else 0
        self.want_draws, self.want_summary = want_draws, want_summary
        self._h = C.c_void_p()
        check(lib().foct_plan_create(kind, batch.array, n_problems, C.byref(spec), C.byref(cfg), int(want_draws),
                                     int(want_summary), C.byref(self._h)))

    def run(self, seed=None):
        check(lib().foct_plan_run(self._h, int(self.cfg.seed if seed is None else seed)))

    def sync(self) -> float:
        ms = C.c_float()
        check(lib().foct_plan_sync(self._h, C.byref(ms)))
        return float(ms.value)

    def timing(self) -> dict:
        a, b = C.c_float(), C.c_float()
        g = [C.c_int() for _ in range(5)]
        check(lib().foct_plan_timing(self._h, C.byref(a), C.byref(b), *[C.byref(v) for v in g]))
        return dict(sample_ms=float(a.value), summary_ms=float(b.value), grid=g[0].value, block=g[1].value,
                    blocks_per_sm=g[2].value, regs=g[3].value, smem_bytes=g[4].value)

    def query(self):
        """Non-blocking: (done, fraction of chain-iterations completed)."""
        d, f = C.c_int(), np.zeros(1)
        check(lib().foct_plan_query(self._h, C.byref(d), abi.as_ptr(f)))
        return bool(d.value), float(f[0])

    def cancel(self):
        check(lib().foct_plan_cancel(self._h))

    def launches(self) -> int:
        return int(lib().foct_plan_launches(self._h))

    def fetch(self):
        out, R = alloc_result(self.kind, self.n, self.Nn, self.cfg, self.want_draws, self.want_summary)
        check(lib().foct_plan_fetch(self._h, C.byref(R)))
        return out

    def close(self):
        if self._h:
            lib().foct_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def monoexp_map(batch: abi.ProblemBatch, n_problems: int, spec: abi.ModelSpec, init=None):
    theta = np.empty((n_problems, 3))
    H = np.empty((n_problems, 3, 3))
    br = np.empty(n_problems)
    st = np.empty(n_problems, dtype=np.int32)
    ip = abi.as_ptr(np.ascontiguousarray(init, dtype=np.float64)) if init is not None else abi.c_double_p()
    check(lib().foct_monoexp_map(batch.array, n_problems, C.byref(spec), ip, abi.as_ptr(theta), abi.as_ptr(H),
                                 abi.as_ptr(br), st.ctypes.data_as(C.POINTER(C.c_int))))
    return theta, H, br, st


def expgp_map(batch: abi.ProblemBatch, n_problems: int, spec: abi.ModelSpec, init=None, hessian=True):
    """fitExpGP(method='optim'): par [n, Nn+7] (theta, yGP, lambda, sigma, br, lp), hessian [n, D, D], status [n]."""
    Nn = batch.array[0].Nn
    D, P_out = abi.dims(abi.FOCT_EXPGP, Nn)
    par = np.empty((n_problems, P_out))
    H = np.empty((n_problems, D, D)) if hessian else None
    st = np.empty(n_problems, dtype=np.int32)
    ip = abi.as_ptr(np.ascontiguousarray(init, dtype=np.float64)) if init is not None else abi.c_double_p()
    check(lib().foct_expgp_map(batch.array, n_problems, C.byref(spec), ip, abi.as_ptr(par), abi.as_ptr(H),
                               st.ctypes.data_as(C.POINTER(C.c_int))))
    return par, H, st


def predict(kind, batch, j, spec, draws):
    draws = np.ascontiguousarray(draws, dtype=np.float64)
    n = draws.shape[0]
    N = batch.array[j].N
    m, r, dl = np.empty((n, N)), np.empty((n, N)), np.empty((n, N))
    check(lib().foct_predict(kind, C.byref(batch.array[j]), C.byref(spec), abi.as_ptr(draws), n, abi.as_ptr(m),
                             abi.as_ptr(r), abi.as_ptr(dl)))
    return m, r, dl


def summary(draws):
    """rstan's summary(fit)$summary for draws[n_draws, chains, n_cols] (or [n_sets, n_draws, chains, n_cols]) on the device."""
    d = np.ascontiguousarray(draws, dtype=np.float64)
    single = d.ndim == 3
    if single:
        d = d[None]
    n_sets, n, c, p = d.shape
    out = np.empty((n_sets, p, abi.FOCT_N_SUMMARY_COLS))
    check(lib().foct_summary(abi.as_ptr(d), n_sets, n, c, p, abi.as_ptr(out)))
    return out[0] if single else out


def fp64_peak(device: int = 0):
    t, f = np.zeros(1), np.zeros(1)
    check(lib().foct_fp64_peak(device, abi.as_ptr(t), abi.as_ptr(f)))
    return float(t[0]), float(f[0])


# ---- the steps either side of the sampling path (include/fitoct_b200.h, MODEL_SPEC §11-13) ----
def _split(packed, Ns):
    offs = np.concatenate([[0], np.cumsum(Ns)]).astype(np.int64)
    return [packed[offs[j]:offs[j + 1]] for j in range(len(Ns))]


def _int_ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def estimate_noise(batch: abi.ProblemBatch, n_problems: int, df: float = 15.0, max_rate: float = 1e4):
    """Returns dict(uy, ySmooth: lists of per-problem arrays; theta [n,2]; info [n,4]; status [n])."""
    Ns = [batch.array[j].N for j in range(n_problems)]
    tot = int(np.sum(Ns))
    uy, ys = np.empty(tot), np.empty(tot)
    th, info = np.empty((n_problems, 2)), np.empty((n_problems, 4))
    st = np.empty(n_problems, dtype=np.int32)
    check(lib().foct_estimate_noise(batch.array, n_problems, float(df), float(max_rate), abi.as_ptr(uy), abi.as_ptr(ys),
                                    abi.as_ptr(th), abi.as_ptr(info), _int_ptr(st)))
    return dict(uy=_split(uy, Ns), ySmooth=_split(ys, Ns), theta=th, info=info, status=st)


def birge_ci(ndf: float) -> np.ndarray:
    ci = np.empty(2)
    check(lib().foct_birge_ci(float(ndf), abi.as_ptr(ci)))
    return ci


def print_br(kind: int, batch: abi.ProblemBatch, n_problems: int, spec: abi.ModelSpec, br):
    """Returns ci [n,2], alert [n] (1 = br outside the 95 % interval: the fit is not OK)."""
    br = np.ascontiguousarray(br, dtype=np.float64)
    ci = np.empty((n_problems, 2))
    alert = np.empty(n_problems, dtype=np.int32)
    check(lib().foct_print_br(kind, batch.array, n_problems, C.byref(spec), abi.as_ptr(br), abi.as_ptr(ci), _int_ptr(alert)))
    return ci, alert


def estimate_exp_prior(batch: abi.ProblemBatch, n_problems: int, priorType: str, theta_map, hessian, ru_theta: float = 0.05):
    """priorType 'mono' | 'abc'.  Returns theta0 [n,3], Sigma0 [n,3,3], ru [n]."""
    if priorType not in ("mono", "abc"):
        raise ValueError("priorType must be 'mono' or 'abc'")
    th = np.ascontiguousarray(theta_map, dtype=np.float64)
    H = np.ascontiguousarray(hessian, dtype=np.float64)
    t0, S0, ru = np.empty((n_problems, 3)), np.empty((n_problems, 3, 3)), np.empty(n_problems)
    check(lib().foct_estimate_exp_prior(batch.array, n_problems, {"mono": 0, "abc": 1}[priorType], abi.as_ptr(th),
                                        abi.as_ptr(H), float(ru_theta), abi.as_ptr(t0), abi.as_ptr(S0), abi.as_ptr(ru)))
    return t0, S0, ru


def pipeline_cfg(**kw) -> abi.PipelineCfg:
    c = abi.PipelineCfg()
    lib().foct_pipeline_cfg_default(C.byref(c))
    for k, v in kw.items():
        if not hasattr(c, k):
            raise TypeError(f"unknown pipeline key {k}")
        setattr(c, k, v)
    return c


def pipeline(batch: abi.ProblemBatch, n_problems: int, pcfg: abi.PipelineCfg, cfg: abi.SamplerCfg, spec_gp=None, draws=False,
             summary=True):
    """foct_pipeline: noise -> MonoExp MAP -> gate -> prior -> fitExpGP on the gated profiles, one call."""
    n = n_problems
    Ns = [batch.array[j].N for j in range(n)]
    tot = int(np.sum(Ns))
    o = dict(uy=np.empty(tot), ySmooth=np.empty(tot), noise_theta=np.empty((n, 2)), mono_theta=np.empty((n, 3)),
             mono_hessian=np.empty((n, 3, 3)), mono_br=np.empty(n), mono_status=np.empty(n, dtype=np.int32),
             br_ci=np.empty((n, 2)), alert=np.empty(n, dtype=np.int32), theta0=np.empty((n, 3)), Sigma0=np.empty((n, 3, 3)),
             ru=np.empty(n), expgp_index=np.full(n, -1, dtype=np.int32), status=np.full(n, -1, dtype=np.int32))
    res, R = alloc_result(abi.FOCT_EXPGP, n, pcfg.Nn, cfg, draws, summary)
    out = abi.PipelineOut()
    for k, v in o.items():
        setattr(out, k, _int_ptr(v) if v.dtype == np.int32 else abi.as_ptr(v))
    out.expgp = R
    check(lib().foct_pipeline(batch.array, n, C.byref(pcfg), C.byref(spec_gp) if spec_gp is not None else None, C.byref(cfg),
                              C.byref(out)))
    k = out.n_expgp
    o["uy"], o["ySmooth"] = _split(o["uy"], Ns), _split(o["ySmooth"], Ns)
    o["n_expgp"] = k
    o["expgp_index"] = o["expgp_index"][:k]
    o["expgp"] = {name: (v[:k] if v is not None else None) for name, v in res.items()}
    return o


# ---- method = 'vb' (MODEL_SPEC §14) ----
def vb(kind: int, batch: abi.ProblemBatch, n_problems: int, spec: abi.ModelSpec, cfg: abi.VbCfg, draws=True):
    """Mean-field ADVI for a batch.  Returns dict(mean [n,P_out], draws [n,output_samples,P_out], mu, omega [n,D], elbo, eta,
    iters, status [n])."""
    Nn = batch.array[0].Nn if kind == abi.FOCT_EXPGP else 0
    out, R = abi.alloc_vb_result(kind, n_problems, Nn, cfg, draws)
    check(lib().foct_vb(kind, batch.array, n_problems, C.byref(spec), C.byref(cfg), C.byref(R)))
    return out


def release_cache():
    """Return the library's cached device buffers to the driver (foct_release_cache)."""
    lib().foct_release_cache()
