"""Host-side mirror of the reference's operator interface for the hot path.

Same names, argument meaning and error behaviour as the calls the reference makes:

    FitOCTLib::fitExpGP(x, y, uy, dataType, Nn, gridType, method, theta0, Sigma0, lambda_rate, rho_scale,
                        nb_warmup, nb_iter, prior_PD, open_progress)            FitOCT.R:110-124, priPost.R:2-16
    FitOCTLib::fitMonoExp(x, y, uy, dataType)                                   FitOCT.R:95

and the same return shapes (`list(fit, method, xGP, prior_PD)`, plotExpGP.R:29-32;
`list(best.theta, cor.theta, fit$par$..., fit$hessian, method)`, FitOCT.R:96-97, plotMonoExp.R:14-16).
The R side of the drop-in (r-pkg/) is a logic-free shim over the same C ABI; R is absent from this image,
so this Python mirror is what the tests and benchmarks drive.  Everything numerical happens in the CUDA
library (fitoct_b200._lib); nothing here falls back to the CPU.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _abi as abi
from . import _lib as L

# FitOCT.R:37-53 defaults, overridden by ctrlParams.yaml (FitOCT.R:55-63)
CTRL_DEFAULTS = dict(
    depthSel=None, dataType=2, subSample=1, smooth_df=15, method="sample", nb_warmup=500, nb_sample=1000,
    modRange=0.5, ru_theta=0.05, lambda_rate=0.1, gridType="internal", Nn=10, rho_scale=0.1, priPost=True,
    priorType="abc",
)


def load_ctrl_params(path: str | None = "ctrlParams.yaml") -> dict:
    """Defaults <- YAML overrides, like FitOCT.R:37-63.  Unknown keys are kept (the R script assigns them too)."""
    import os

    pars = dict(CTRL_DEFAULTS)
    if path and os.path.exists(path):
        import yaml

        with open(path) as fh:
            loaded = yaml.safe_load(fh) or {}
        pars.update(loaded)
    return pars


def resolve_rho(rho_scale: float, Nn: int) -> float:
    """`ifelse(rho_scale==0, 1./Nn, rho_scale)` — FitOCT.R:119, priPost.R:11, server.R:420-422."""
    return 1.0 / Nn if rho_scale == 0 else float(rho_scale)


@dataclass
class StanFit:
    """stanfit-like result (SURVEY a-11): what plotExpGP.R:9-11,41-47 and server.R:88-104 read."""

    par_names: list
    draws: np.ndarray           # [n_saved, chains, P_out], warm-up first when saved
    sampler_params: np.ndarray  # [n_saved, chains, 6]
    n_warmup: int
    n_iter: int
    save_warmup: bool
    summary_table: np.ndarray   # [P_out, 11] over post-warm-up draws (computed on the device)
    stepsize: np.ndarray
    inv_metric: np.ndarray
    n_divergent: np.ndarray

    def _post(self):
        return self.draws[self.n_warmup:] if self.save_warmup else self.draws

    def _cols(self, pars):
        if pars is None:
            return list(range(len(self.par_names)))
        if isinstance(pars, str):
            pars = [pars]
        idx = []
        for p in pars:
            hit = [i for i, nme in enumerate(self.par_names) if nme == p or nme.startswith(p + "[")]
            if not hit:
                raise KeyError(f"no parameter {p!r}; have {self.par_names}")
            idx += hit
        return idx

    def extract(self, pars=None, inc_warmup=False) -> dict:
        """rstan::extract: post-warm-up draws, chains merged (plotExpGP.R:11)."""
        d = self.draws if (inc_warmup and self.save_warmup) else self._post()
        flat = d.reshape(-1, d.shape[-1])
        groups: dict = {}
        for i in self._cols(pars):
            groups.setdefault(self.par_names[i].split("[")[0], []).append(i)
        return {k: (flat[:, v[0]] if len(v) == 1 and "[" not in self.par_names[v[0]] else flat[:, v]) for k, v in groups.items()}

    def as_matrix(self, pars=None) -> np.ndarray:
        """as.matrix(fit, pars=...) — plotExpGP.R:44."""
        d = self._post()
        return d.reshape(-1, d.shape[-1])[:, self._cols(pars)]

    def summary(self, pars=None) -> dict:
        """rstan::summary(fit)$summary with the rstan column names (server.R:88-104)."""
        idx = self._cols(pars)
        return {"rownames": [self.par_names[i] for i in idx], "colnames": list(abi.SUMMARY_COL_NAMES),
                "summary": self.summary_table[idx]}

    def __str__(self):
        s = self.summary()
        head = f"{'':12s}" + "".join(f"{c:>10s}" for c in s["colnames"])
        rows = [f"{n:12s}" + "".join(f"{v:10.4g}" for v in r) for n, r in zip(s["rownames"], s["summary"])]
        return "\n".join([f"Inference for fitoct_b200 model: {self.draws.shape[1]} chains, iter={self.n_iter}, "
                          f"warmup={self.n_warmup}", head] + rows)


def monitor(draws, par_names=None) -> dict:
    """rstan::monitor(sims) / summary(fit)$summary for draws the caller holds: an [iterations, chains, parameters] array
    (the shape of as.array(stanfit)), post-warm-up.  Computed on the device by the kernel that summarises a fit
    (foct_summary); the table a merged or thinned stanfit needs again (server.R:88-104)."""
    d = np.asarray(draws, dtype=np.float64)
    if d.ndim != 3:
        raise ValueError("draws must be [iterations, chains, parameters]")
    names = list(par_names) if par_names is not None else [f"V{i + 1}" for i in range(d.shape[2])]
    if len(names) != d.shape[2]:
        raise ValueError("one name per parameter")
    return {"rownames": names, "colnames": list(abi.SUMMARY_COL_NAMES), "summary": L.summary(d)}


def _grid_code(gridType):
    if isinstance(gridType, str):
        if gridType not in ("internal", "extremal"):
            raise ValueError("gridType must be 'internal' or 'extremal' (ShinyInterface/ui.R:211-218)")
        return 0 if gridType == "internal" else 1
    return int(gridType)


def _one_problem(x, y, uy, dataType, Nn, gridType, theta0, Sigma0, lambda_rate, rho, prior_PD, pid=0):
    return dict(x=x, y=y, uy=uy, dataType=dataType, Nn=Nn, gridType=_grid_code(gridType), rho=rho,
                lambda_rate=lambda_rate, theta0=theta0, Sigma0=Sigma0, prior_PD=prior_PD, id=pid)


def _stanfit_from(out, j, kind, Nn, cfg):
    return StanFit(par_names=abi.param_names(kind, Nn), draws=out["draws"][j], sampler_params=out["sampler_params"][j],
                   n_warmup=cfg.n_warmup, n_iter=cfg.n_iter, save_warmup=bool(cfg.save_warmup),
                   summary_table=out["summary"][j], stepsize=out["stepsize"][j], inv_metric=out["inv_metric"][j],
                   n_divergent=out["n_divergent"][j])


def fitExpGP(x, y, uy, dataType=2, Nn=10, gridType="internal", method="sample", theta0=None, Sigma0=None,
             lambda_rate=0.1, rho_scale=0.1, nb_warmup=500, nb_iter=1500, prior_PD=0, open_progress=False, *,
             chains=4, seed=1234, spec=None, control=None, init=None):
    """Drop-in for FitOCTLib::fitExpGP (FitOCT.R:110-124).  Returns dict(fit, method, xGP, prior_PD).

    `rho_scale` is taken as already resolved by the caller when non-zero; 0 is resolved to 1/Nn here too, so
    both calling styles of the reference work.  `control` accepts rstan's adapt_delta / max_treedepth.
    """
    if theta0 is None or Sigma0 is None:
        raise ValueError("theta0 and Sigma0 are required (estimateExpPrior output, FitOCT.R:103-107)")
    if method not in ("sample", "optim", "vb"):
        raise ValueError("method must be one of 'sample', 'optim', 'vb' (FitOCT.R:42)")
    spec = spec or abi.default_spec(abi.FOCT_EXPGP)
    control = control or {}
    if method == "vb":
        # Stan's mean-field ADVI (MODEL_SPEC §14).  rstan::vb's arguments pass through `control`.  omega0: Stan starts the
        # approximation at unit standard deviations (omega = 0), where this model's draws have 1 + dL <= 0 and ADVI stops with
        # "dropped evaluations"; the mirror starts narrower unless told otherwise (control = dict(omega0 = 0) is Stan's).
        vcfg = abi.default_vb_cfg(seed=int(seed), omega0=float(control.get("omega0", -3.0)),
                                  **{k: control[k] for k in ("iter", "grad_samples", "elbo_samples", "eval_elbo", "output_samples",
                                                             "adapt_engaged", "adapt_iter", "eta", "tol_rel_obj") if k in control})
        keep = None
        if init is not None:
            keep = np.ascontiguousarray(init, dtype=np.float64).reshape(1, Nn + 5)
            vcfg.init_mode = 2
            vcfg.init = abi.as_ptr(keep)
        batch = abi.make_problems([_one_problem(x, y, uy, dataType, Nn, gridType, theta0, Sigma0, lambda_rate,
                                                resolve_rho(rho_scale, Nn), prior_PD)])
        o = L.vb(abi.FOCT_EXPGP, batch, 1, spec, vcfg, draws=True)
        if o["status"][0] == 2:
            raise RuntimeError("fitExpGP(method='vb'): ADVI failed (non-finite gradient / ELBO, or all proposed step sizes "
                               "failed) — Stan reports the same as 'dropped evaluations'")
        ns = vcfg.output_samples
        d = o["draws"][0]
        # rstan's table for a vb fit: independent draws, so se_mean = sd/sqrt(n), n_eff = n, no Rhat
        tab = np.column_stack([d.mean(0), d.std(0, ddof=1) / np.sqrt(ns), d.std(0, ddof=1),
                               *np.quantile(d, [0.025, 0.25, 0.5, 0.75, 0.975], axis=0), np.full(d.shape[1], float(ns)),
                               np.full(d.shape[1], np.nan), np.full(d.shape[1], float(ns))])
        fit = StanFit(par_names=abi.param_names(abi.FOCT_EXPGP, Nn), draws=o["draws"][0][:, None, :],
                      sampler_params=np.zeros((ns, 1, 6)), n_warmup=0, n_iter=ns, save_warmup=False,
                      summary_table=tab, stepsize=np.array([o["eta"][0]]), inv_metric=np.exp(2 * o["omega"][0])[None, :],
                      n_divergent=np.zeros(1))
        fit.vb = dict(mean=o["mean"][0], mu=o["mu"][0], omega=o["omega"][0], elbo=float(o["elbo"][0]), eta=float(o["eta"][0]),
                      iterations=int(o["iters"][0]), converged=bool(o["status"][0] == 0))
        return dict(fit=fit, method=method, xGP=L.grid(Nn, _grid_code(gridType)), prior_PD=prior_PD)
    if method == "optim":
        # MAP + Hessian (MODEL_SPEC §10): the Shiny default (ui.R:107-114); consumers plotExpGP.R:13-17, server.R:164-173
        batch = abi.make_problems([_one_problem(x, y, uy, dataType, Nn, gridType, theta0, Sigma0, lambda_rate,
                                                resolve_rho(rho_scale, Nn), prior_PD)])
        q0 = None if init is None else np.ascontiguousarray(init, dtype=np.float64).reshape(1, Nn + 5)
        par, H, st = L.expgp_map(batch, 1, spec, init=q0, hessian=True)
        row = par[0]
        m, resid, dl = L.predict(abi.FOCT_EXPGP, batch, 0, spec, row[None, :])
        fit = dict(par=dict(theta=row[:3], yGP=row[3:3 + Nn], **{"lambda": row[3 + Nn]}, sigma=row[4 + Nn], br=row[5 + Nn],
                            m=m[0], resid=resid[0], dL=dl[0]), value=row[6 + Nn], hessian=H[0], return_code=int(st[0]))
        return dict(fit=fit, method=method, xGP=L.grid(Nn, _grid_code(gridType)), prior_PD=prior_PD)
    cfg = abi.default_cfg(chains=chains, n_warmup=int(nb_warmup), n_iter=int(nb_iter), seed=int(seed), save_warmup=1,
                          adapt_delta=float(control.get("adapt_delta", 0.8)),
                          max_treedepth=int(control.get("max_treedepth", 10)))
    keep = None
    if init is not None:
        keep = np.ascontiguousarray(init, dtype=np.float64).reshape(chains, Nn + 5)
        cfg.init_mode = 2
        cfg.init = abi.as_ptr(keep)
    batch = abi.make_problems([_one_problem(x, y, uy, dataType, Nn, gridType, theta0, Sigma0, lambda_rate,
                                            resolve_rho(rho_scale, Nn), prior_PD)])
    out = L.sample(abi.FOCT_EXPGP, batch, 1, spec, cfg, draws=True, summary=True)
    fit = _stanfit_from(out, 0, abi.FOCT_EXPGP, Nn, cfg)
    return dict(fit=fit, method=method, xGP=L.grid(Nn, _grid_code(gridType)), prior_PD=prior_PD)


def fitExpGP_batch(x, Y, UY, theta0, Sigma0, *, dataType=2, Nn=10, gridType="internal", lambda_rate=0.1, rho_scale=0.1,
                   nb_warmup=500, nb_iter=1500, prior_PD=0, chains=4, seed=1234, spec=None, control=None,
                   devices=None, ids=None, keep_draws=False):
    """Batch form of fitExpGP: one call for a whole directory of profiles, sharded over
    `devices` as independent shards.  Returns the device-computed summaries (and the draws on request)."""
    spec = spec or abi.default_spec(abi.FOCT_EXPGP)
    control = control or {}
    cfg = abi.default_cfg(chains=chains, n_warmup=int(nb_warmup), n_iter=int(nb_iter), seed=int(seed),
                          adapt_delta=float(control.get("adapt_delta", 0.8)),
                          max_treedepth=int(control.get("max_treedepth", 10)))
    batch = abi.make_problems_dense(x, Y, UY, theta0, Sigma0, dataType=dataType, Nn=Nn, gridType=_grid_code(gridType),
                                    rho=resolve_rho(rho_scale, Nn), lambda_rate=lambda_rate, prior_PD=prior_PD, ids=ids)
    n = np.asarray(Y).shape[0]
    out = L.sample(abi.FOCT_EXPGP, batch, n, spec, cfg, draws=keep_draws, summary=True, devices=devices)
    out["par_names"] = abi.param_names(abi.FOCT_EXPGP, Nn)
    out["xGP"] = L.grid(Nn, _grid_code(gridType))
    return out


def fitMonoExp(x, y, uy, dataType=2, method="optim", *, nb_warmup=500, nb_iter=1500, chains=4, seed=1234, spec=None):
    """Drop-in for FitOCTLib::fitMonoExp (FitOCT.R:95).  method='optim' is what the reference uses (MAP +
    Hessian); method='sample' is BASELINE.json configs[0] (4-chain NUTS, D = 3)."""
    spec = spec or abi.default_spec(abi.FOCT_MONOEXP)
    batch = abi.make_problems([_one_problem(x, y, uy, dataType, 0, 0, (0, 0, 1), np.eye(3), 0.0, 1.0, 0)])
    theta, H, br, st = L.monoexp_map(batch, 1, spec)
    best = theta[0]
    cov = np.linalg.inv(-H[0])
    sd = np.sqrt(np.diag(cov))
    cor = cov / np.outer(sd, sd)
    if method == "optim":
        row = np.array([[best[0], best[1], best[2], br[0], 0.0]])
        m, resid, _ = L.predict(abi.FOCT_MONOEXP, batch, 0, spec, row)
        fit = dict(par=dict(theta=best, m=m[0], resid=resid[0], br=br[0]), hessian=H[0], return_code=int(st[0]))
        return dict(best_theta=best, cor_theta=cor, fit=fit, method="optim")
    if method != "sample":
        raise ValueError("method must be 'optim' or 'sample'")
    cfg = abi.default_cfg(chains=chains, n_warmup=int(nb_warmup), n_iter=int(nb_iter), seed=int(seed), save_warmup=1)
    init = np.ascontiguousarray(np.tile(best, (chains, 1)))
    cfg.init_mode = 2
    cfg.init = abi.as_ptr(init)
    out = L.sample(abi.FOCT_MONOEXP, batch, 1, spec, cfg, draws=True, summary=True)
    fit = _stanfit_from(out, 0, abi.FOCT_MONOEXP, 0, cfg)
    return dict(best_theta=fit.summary_table[:3, 0], cor_theta=cor, fit=fit, method="sample")


# ---- the steps either side of the path (MODEL_SPEC §11-13) ----
def _xy_problem(x, y, dataType=2):
    x = np.asarray(x, dtype=np.float64)
    return dict(x=x, y=y, uy=np.ones_like(x), dataType=dataType, Nn=0, gridType=0, rho=1.0, lambda_rate=0.0,
                theta0=(0.0, 0.0, 1.0), Sigma0=np.eye(3), prior_PD=0)


def estimateNoise(x, y, df=15, maxRate=10000):
    """Drop-in for FitOCTLib::estimateNoise (FitOCT.R:89-91, server.R:309-311): `uy`, `ySmooth`, `theta` (= a_1, a_2 of
    uy = a_1 exp(-x/a_2), plotNoise.R:4-6), plus `fit` (spar, lambda, df of the smoother) and `method`."""
    batch = abi.make_problems([_xy_problem(x, y)])
    o = L.estimate_noise(batch, 1, df=float(df), max_rate=float(maxRate))
    if o["status"][0] == 3:
        raise ValueError("estimateNoise: x must be strictly increasing")
    info = o["info"][0]
    return dict(uy=o["uy"][0], ySmooth=o["ySmooth"][0], theta=o["theta"][0], method="optim",
                fit=dict(spar=info[0], **{"lambda": info[1]}, df=info[2], iterations=int(info[3]), return_code=int(o["status"][0])))


def printBr(fit, N=None, n_par=None, silent=False):
    """Drop-in for FitOCTLib::printBr (plotMonoExp.R:10, plotExpGP.R:22, server.R:354,379,499,510).  `fit` is what
    fitMonoExp / fitExpGP return in `$fit`: an optimizing-style dict (`par$br`, `par$resid`) or a StanFit (posterior mean
    of `br`).  Returns dict(br, CI95, alert) with alert None when the fit is OK (the test at FitOCT.R:100)."""
    if isinstance(fit, StanFit):
        d = fit.extract(["br"])
        br = float(np.mean(d["br"]))
        n_par = n_par if n_par is not None else len([p for p in fit.par_names if p.startswith(("theta", "yGP"))])
        if N is None:
            raise ValueError("printBr on a sampled fit needs N (the stanfit does not carry the data)")
    else:
        br = float(fit["par"]["br"])
        N = N if N is not None else len(fit["par"]["resid"])
        n_par = n_par if n_par is not None else len(np.atleast_1d(fit["par"]["theta"])) + len(np.atleast_1d(fit["par"].get("yGP", [])))
    ndf = N - n_par
    ci = L.birge_ci(ndf)
    ok = ci[0] <= br <= ci[1]
    alert = None if ok else "!!! WARNING !!! br out of interval"
    if not silent:
        print(f"br   : {br:.3g}\nCI95 : [{ci[0]:.3g}, {ci[1]:.3g}]" + ("" if ok else f"\n{alert}"))
    return dict(br=br, CI95=ci, alert=alert)


def estimateExpPrior(x, uy, dataType, priorType="mono", out=None, ru_theta=0.05, eps=1e-3):
    """Drop-in for FitOCTLib::estimateExpPrior (FitOCT.R:103-107, server.R:396-404).  `out` is fitMonoExp's result
    (method 'optim'); returns dict(theta0, Sigma0[, ru]).  `eps` (the ABC tolerance) has no role in the closed-form
    matching of MODEL_SPEC §13 and is accepted for signature compatibility."""
    if out is None:
        raise ValueError("out= (the fitMonoExp result) is required")
    fit = out["fit"]
    x = np.asarray(x, dtype=np.float64)
    y = fit["par"]["m"] + fit["par"]["resid"]
    p = _xy_problem(x, y, dataType)
    p["uy"] = np.asarray(uy, dtype=np.float64)
    batch = abi.make_problems([p])
    t0, S0, ru = L.estimate_exp_prior(batch, 1, priorType, np.asarray(out["best_theta"])[None, :], np.asarray(fit["hessian"])[None],
                                      ru_theta=float(ru_theta))
    return dict(theta0=t0[0], Sigma0=S0[0], ru=float(ru[0]))


def FitOCT_batch(x, Y, ctrl=None, *, chains=4, seed=1234, gate=True, keep_draws=False, spec=None):
    """FitOCT.R:74-124 for a batch of profiles on a shared depth grid: estimateNoise -> fitMonoExp -> printBr gate ->
    estimateExpPrior -> fitExpGP(method = ctrl['method']) on the profiles the gate lets through.  `ctrl` holds
    ctrlParams.yaml keys (load_ctrl_params()).  method 'sample' is ONE library call (foct_pipeline); 'optim' and 'vb' chain
    the batched entry points (foct_estimate_noise, foct_monoexp_map, foct_print_br, foct_estimate_exp_prior, then
    foct_expgp_map / foct_vb on the gated subset).  NB the reference `break`s out of its dataset loop at the first profile
    whose MonoExp fit is OK (FitOCT.R:100); a batch treats that as `next`."""
    c = dict(CTRL_DEFAULTS)
    c.update(ctrl or {})
    from . import io as fio

    method = c.get("method", "sample")
    if method not in ("sample", "optim", "vb"):
        raise ValueError("method must be one of 'sample', 'optim', 'vb' (FitOCT.R:42)")
    Y = np.atleast_2d(np.asarray(Y, dtype=np.float64))
    n = Y.shape[0]
    xs, Ys = [], []
    for j in range(n):
        xj, yj = fio.selX(x, Y[j], c.get("depthSel"), c.get("subSample", 1))
        xs.append(xj)
        Ys.append(yj)
    Ysel = np.stack(Ys)
    Nn, gcode = int(c["Nn"]), _grid_code(c["gridType"])
    th00, eye = np.tile([0.0, 0.0, 1.0], (n, 1)), np.tile(np.eye(3), (n, 1, 1))
    batch = abi.make_problems_dense(xs[0], Ysel, np.ones_like(Ysel), th00, eye, dataType=int(c["dataType"]), Nn=0)
    if method == "sample":
        pc = L.pipeline_cfg(smooth_df=float(c["smooth_df"]), prior_type={"mono": 0, "abc": 1}[c["priorType"]],
                            ru_theta=float(c["ru_theta"]), Nn=Nn, gridType=gcode, rho_scale=float(c["rho_scale"]),
                            lambda_rate=float(c["lambda_rate"]), gate=int(bool(gate)))
        cfg = abi.default_cfg(chains=chains, n_warmup=int(c["nb_warmup"]), n_iter=int(c["nb_warmup"]) + int(c["nb_sample"]),
                              seed=int(seed))
        out = L.pipeline(batch, n, pc, cfg, spec_gp=spec, draws=keep_draws, summary=True)
        out["sampler_cfg"] = cfg
    else:
        nz = L.estimate_noise(batch, n, df=float(c["smooth_df"]))
        UY = np.stack(nz["uy"])
        spec_m = abi.default_spec(abi.FOCT_MONOEXP)
        mono = abi.make_problems_dense(xs[0], Ysel, UY, th00, eye, dataType=int(c["dataType"]), Nn=0)
        th, H, br, st = L.monoexp_map(mono, n, spec_m)
        ci, alert = L.print_br(abi.FOCT_MONOEXP, mono, n, spec_m, br)
        t0, S0, ru = L.estimate_exp_prior(mono, n, c["priorType"], th, H, ru_theta=float(c["ru_theta"]))
        idx = np.flatnonzero(alert) if gate else np.arange(n)
        out = dict(uy=nz["uy"], ySmooth=nz["ySmooth"], noise_theta=nz["theta"], mono_theta=th, mono_hessian=H, mono_br=br,
                   mono_status=st, br_ci=ci, alert=alert, theta0=t0, Sigma0=S0, ru=ru, n_expgp=int(idx.size),
                   expgp_index=idx.astype(np.int32), expgp=None)
        if idx.size:
            gp = abi.make_problems_dense(xs[0], Ysel[idx], UY[idx], t0[idx], S0[idx], dataType=int(c["dataType"]), Nn=Nn,
                                         gridType=gcode, rho=resolve_rho(float(c["rho_scale"]), Nn),
                                         lambda_rate=float(c["lambda_rate"]), ids=idx)
            sg = spec or abi.default_spec(abi.FOCT_EXPGP)
            if method == "optim":
                par, Hq, stq = L.expgp_map(gp, int(idx.size), sg, hessian=True)
                out["expgp"] = dict(par=par, hessian=Hq, status=stq)
            else:
                vc = abi.default_vb_cfg(seed=int(seed), omega0=float(c.get("vb_omega0", -3.0)))
                out["expgp"] = L.vb(abi.FOCT_EXPGP, gp, int(idx.size), sg, vc, draws=keep_draws)
    out["method"] = method
    out["x"] = xs[0]
    out["par_names"] = abi.param_names(abi.FOCT_EXPGP, Nn)
    out["xGP"] = L.grid(Nn, gcode)
    return out
