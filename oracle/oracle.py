"""ctypes wrapper of the CPU oracle (oracle/libfoct_oracle.so).  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED — see oracle/foct_oracle.h.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from fitoct_b200 import _abi as abi  # noqa: E402  (struct definitions only)

_LIB = None
_FAST = None


def fast_lib():
    """The TIMING build (-O3, AVX2 + FMA; oracle/Makefile): same entry points, used only by bench.py's CPU legs."""
    global _FAST
    if _FAST is None:
        path = os.path.join(_HERE, "libfoct_oracle_fast.so")
        if not os.path.exists(path):
            build()
        _FAST = C.CDLL(path)
        _FAST.foct_oracle_sample.argtypes = [C.c_int, C.POINTER(abi.Problem), C.c_int, C.POINTER(abi.ModelSpec),
                                             C.POINTER(abi.SamplerCfg), C.POINTER(abi.Result), C.c_int]
    return _FAST


def build() -> str:
    path = os.path.join(_HERE, "libfoct_oracle.so")
    subprocess.run(["make", "-s", "-C", _HERE], check=True)
    return path


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libfoct_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        dp, ip = abi.c_double_p, C.POINTER(C.c_int)
        L.foct_oracle_philox.argtypes = [C.POINTER(C.c_uint32)] * 3
        L.foct_oracle_grid.argtypes = [C.c_int, C.c_int, dp]
        L.foct_oracle_basis.argtypes = [C.POINTER(abi.Problem), C.POINTER(abi.ModelSpec), dp]
        L.foct_oracle_logp_grad.argtypes = [C.c_int, C.POINTER(abi.Problem), C.POINTER(abi.ModelSpec), dp, dp,
                                            C.c_int, dp, dp, dp, dp]
        L.foct_oracle_sample.argtypes = [C.c_int, C.POINTER(abi.Problem), C.c_int, C.POINTER(abi.ModelSpec),
                                         C.POINTER(abi.SamplerCfg), C.POINTER(abi.Result), C.c_int]
        L.foct_oracle_sample_analytic.argtypes = [C.c_int, C.c_int, dp, C.POINTER(abi.SamplerCfg), dp, dp]
        L.foct_oracle_summary.argtypes = [dp, C.c_int, C.c_int, C.c_int, dp]
        L.foct_oracle_monoexp_map.argtypes = [C.POINTER(abi.Problem), C.c_int, C.POINTER(abi.ModelSpec), dp, dp, dp,
                                              dp, ip]
        L.foct_oracle_expgp_map.argtypes = [C.POINTER(abi.Problem), C.c_int, C.POINTER(abi.ModelSpec), dp, dp, dp, ip]
        L.foct_oracle_predict.argtypes = [C.c_int, C.POINTER(abi.Problem), C.POINTER(abi.ModelSpec), dp, C.c_int,
                                          dp, dp, dp]
        L.foct_oracle_nknots.argtypes = [C.c_int]
        L.foct_oracle_smooth_spline.argtypes = [C.c_int, dp, dp, C.c_double, C.c_int, C.c_double, dp, dp]
        L.foct_oracle_noise_fit.argtypes = [C.c_int, dp, dp, C.c_double, dp]
        L.foct_oracle_estimate_noise.argtypes = [C.POINTER(abi.Problem), C.c_int, C.c_double, C.c_double, dp, dp, dp, dp]
        L.foct_oracle_qchisq.argtypes = [C.c_double, C.c_double]
        L.foct_oracle_qchisq.restype = C.c_double
        L.foct_oracle_print_br.argtypes = [dp, C.c_int, C.c_double, dp, ip]
        L.foct_oracle_exp_prior.argtypes = [C.POINTER(abi.Problem), C.c_int, C.c_int, dp, dp, C.c_double, dp, dp, dp]
        L.foct_oracle_vb.argtypes = [C.c_int, C.POINTER(abi.Problem), C.c_int, C.POINTER(abi.ModelSpec), C.POINTER(abi.VbCfg),
                                     C.POINTER(abi.VbResult)]
        L.foct_oracle_vb_analytic.argtypes = [C.c_int, C.c_int, dp, C.POINTER(abi.VbCfg), dp, dp, dp, dp]
        L.foct_oracle_normal.argtypes = [C.POINTER(C.c_uint32)]
        L.foct_oracle_normal.restype = C.c_double
        _LIB = L
    return _LIB


def _check(rc, what):
    if rc < 0:
        raise RuntimeError(f"{what} failed with code {rc}")
    return rc


def philox(counter, key):
    c = (C.c_uint32 * 4)(*counter)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().foct_oracle_philox(c, k, o)
    return [int(v) for v in o]


def grid(Nn, gridType):
    out = np.empty(Nn)
    _check(lib().foct_oracle_grid(Nn, gridType, abi.as_ptr(out)), "grid")
    return out


def basis(batch: abi.ProblemBatch, j: int, spec: abi.ModelSpec):
    p = batch.array[j]
    B = np.empty((p.Nn, p.N))
    _check(lib().foct_oracle_basis(C.byref(p), C.byref(spec), abi.as_ptr(B)), "basis")
    return B


def logp_grad(kind, batch: abi.ProblemBatch, j: int, spec, q, B=None, want_abs=False):
    """q: [n_q, D].  Returns lp[n_q], grad[n_q,D], chi2[n_q] (, abs_terms[n_q,D])."""
    q = np.ascontiguousarray(q, dtype=np.float64)
    n_q, D = q.shape
    lp = np.empty(n_q)
    g = np.empty((n_q, D))
    chi2 = np.empty(n_q)
    at = np.empty((n_q, D)) if want_abs else None
    Bp = abi.as_ptr(np.ascontiguousarray(B)) if B is not None else abi.c_double_p()
    _check(lib().foct_oracle_logp_grad(kind, C.byref(batch.array[j]), C.byref(spec), Bp, abi.as_ptr(q), n_q,
                                       abi.as_ptr(lp), abi.as_ptr(g), abi.as_ptr(chi2), abi.as_ptr(at)), "logp_grad")
    return (lp, g, chi2, at) if want_abs else (lp, g, chi2)


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
    )
    R = abi.Result()
    for k, v in out.items():
        setattr(R, k, abi.as_ptr(v))
    return out, R


def sample(kind, batch: abi.ProblemBatch, n_problems, spec, cfg, draws=True, summary=True, n_threads=0, fast=False):
    Nn = batch.array[0].Nn if kind == abi.FOCT_EXPGP else 0
    out, R = alloc_result(kind, n_problems, Nn, cfg, draws, summary)
    L = fast_lib() if fast else lib()
    used = _check(L.foct_oracle_sample(kind, batch.array, n_problems, C.byref(spec), C.byref(cfg), C.byref(R),
                                       n_threads), "sample")
    out["threads"] = used
    return out


def sample_analytic(target, par, cfg):
    par = np.ascontiguousarray(par, dtype=np.float64)
    D = par.size if target == 0 else (int(round(par.size ** 0.5)) if target == 2 else 1)
    n_saved = cfg.n_iter if cfg.save_warmup else cfg.n_iter - cfg.n_warmup
    draws = np.empty((n_saved, cfg.chains, D))
    sp = np.empty((n_saved, cfg.chains, 6))
    _check(lib().foct_oracle_sample_analytic(target, D, abi.as_ptr(par), C.byref(cfg), abi.as_ptr(draws),
                                             abi.as_ptr(sp)), "sample_analytic")
    return draws, sp


def summary(draws):
    """draws: [n, chains, P] -> [P, 11]."""
    draws = np.ascontiguousarray(draws, dtype=np.float64)
    n, Cn, P = draws.shape
    out = np.empty((P, abi.FOCT_N_SUMMARY_COLS))
    _check(lib().foct_oracle_summary(abi.as_ptr(draws), n, Cn, P, abi.as_ptr(out)), "summary")
    return out


def monoexp_map(batch: abi.ProblemBatch, n_problems, spec, init=None):
    theta = np.empty((n_problems, 3))
    H = np.empty((n_problems, 3, 3))
    br = np.empty(n_problems)
    st = np.empty(n_problems, dtype=np.int32)
    ip = abi.as_ptr(np.ascontiguousarray(init, dtype=np.float64)) if init is not None else abi.c_double_p()
    _check(lib().foct_oracle_monoexp_map(batch.array, n_problems, C.byref(spec), ip, abi.as_ptr(theta), abi.as_ptr(H),
                                         abi.as_ptr(br), st.ctypes.data_as(C.POINTER(C.c_int))), "monoexp_map")
    return theta, H, br, st


def expgp_map(batch: abi.ProblemBatch, n_problems, spec, init=None, hessian=True):
    Nn = batch.array[0].Nn
    D, P_out = abi.dims(abi.FOCT_EXPGP, Nn)
    par = np.empty((n_problems, P_out))
    H = np.empty((n_problems, D, D)) if hessian else None
    st = np.empty(n_problems, dtype=np.int32)
    ip = abi.as_ptr(np.ascontiguousarray(init, dtype=np.float64)) if init is not None else abi.c_double_p()
    _check(lib().foct_oracle_expgp_map(batch.array, n_problems, C.byref(spec), ip, abi.as_ptr(par), abi.as_ptr(H),
                                       st.ctypes.data_as(C.POINTER(C.c_int))), "expgp_map")
    return par, H, st


def predict(kind, batch, j, spec, draws):
    draws = np.ascontiguousarray(draws, dtype=np.float64)
    n = draws.shape[0]
    N = batch.array[j].N
    m, r, dl = np.empty((n, N)), np.empty((n, N)), np.empty((n, N))
    _check(lib().foct_oracle_predict(kind, C.byref(batch.array[j]), C.byref(spec), abi.as_ptr(draws), n, abi.as_ptr(m),
                                     abi.as_ptr(r), abi.as_ptr(dl)), "predict")
    return m, r, dl


# ---- steps either side of the path (MODEL_SPEC §11-13) ----
def nknots(n):
    return lib().foct_oracle_nknots(int(n))


def smooth_spline(x, y, df=15.0, all_knots=False, spar=None):
    """R-style smooth.spline(x, y, df) (or at a fixed spar).  Returns ySmooth, dict(spar, lambda, df, evals)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    out = np.empty_like(x)
    info = np.empty(4)
    _check(lib().foct_oracle_smooth_spline(x.size, abi.as_ptr(x), abi.as_ptr(y), float(df), int(all_knots),
                                           float("nan") if spar is None else float(spar), abi.as_ptr(out),
                                           abi.as_ptr(info)), "smooth_spline")
    return out, dict(spar=info[0], lam=info[1], df=info[2], evals=int(info[3]))


def noise_fit(x, resid, max_rate=1e4):
    x = np.ascontiguousarray(x, dtype=np.float64)
    r = np.ascontiguousarray(resid, dtype=np.float64)
    th = np.empty(2)
    _check(lib().foct_oracle_noise_fit(x.size, abi.as_ptr(x), abi.as_ptr(r), float(max_rate), abi.as_ptr(th)), "noise_fit")
    return th


def estimate_noise(batch: abi.ProblemBatch, n, df=15.0, max_rate=1e4):
    """Returns uy, ySmooth (lists of per-problem arrays), theta [n,2], info [n,4]."""
    Ns = [batch.array[j].N for j in range(n)]
    tot = int(np.sum(Ns))
    uy, ys = np.empty(tot), np.empty(tot)
    th, info = np.empty((n, 2)), np.empty((n, 4))
    _check(lib().foct_oracle_estimate_noise(batch.array, n, float(df), float(max_rate), abi.as_ptr(uy), abi.as_ptr(ys),
                                            abi.as_ptr(th), abi.as_ptr(info)), "estimate_noise")
    offs = np.concatenate([[0], np.cumsum(Ns)])
    return ([uy[offs[j]:offs[j + 1]] for j in range(n)], [ys[offs[j]:offs[j + 1]] for j in range(n)], th, info)


def qchisq(p, ndf):
    return lib().foct_oracle_qchisq(float(p), float(ndf))


def print_br(br, ndf):
    br = np.ascontiguousarray(br, dtype=np.float64)
    ci = np.empty(2)
    alert = np.empty(br.size, dtype=np.int32)
    _check(lib().foct_oracle_print_br(abi.as_ptr(br), br.size, float(ndf), abi.as_ptr(ci),
                                      alert.ctypes.data_as(C.POINTER(C.c_int))), "print_br")
    return ci, alert


def exp_prior(batch: abi.ProblemBatch, n, priorType, theta_map, hessian, ru_theta=0.05):
    """priorType 'mono' | 'abc'.  Returns theta0 [n,3], Sigma0 [n,3,3], ru [n]."""
    th = np.ascontiguousarray(theta_map, dtype=np.float64)
    H = np.ascontiguousarray(hessian, dtype=np.float64)
    t0, S0, ru = np.empty((n, 3)), np.empty((n, 3, 3)), np.empty(n)
    _check(lib().foct_oracle_exp_prior(batch.array, n, {"mono": 0, "abc": 1}[priorType], abi.as_ptr(th), abi.as_ptr(H),
                                       float(ru_theta), abi.as_ptr(t0), abi.as_ptr(S0), abi.as_ptr(ru)), "exp_prior")
    return t0, S0, ru


# ---- method = 'vb' (MODEL_SPEC §14) ----
def vb(kind, batch: abi.ProblemBatch, n, spec, cfg: abi.VbCfg, draws=True):
    Nn = batch.array[0].Nn if kind == abi.FOCT_EXPGP else 0
    out, R = abi.alloc_vb_result(kind, n, Nn, cfg, draws)
    _check(lib().foct_oracle_vb(kind, batch.array, n, C.byref(spec), C.byref(cfg), C.byref(R)), "vb")
    return out


def vb_analytic(target, par, cfg: abi.VbCfg, q0):
    par = np.ascontiguousarray(par, dtype=np.float64)
    q0 = np.ascontiguousarray(q0, dtype=np.float64)
    D = q0.size
    mu, om, info = np.empty(D), np.empty(D), np.empty(3)
    st = lib().foct_oracle_vb_analytic(target, D, abi.as_ptr(par), C.byref(cfg), abi.as_ptr(q0), abi.as_ptr(mu), abi.as_ptr(om),
                                       abi.as_ptr(info))
    return dict(mu=mu, omega=om, elbo=info[0], eta=info[1], iters=int(info[2]), status=st)
