"""On-disk formats either side of the hot path.

* `Courbe.csv` reader — what FitOCT.R:84 reads (`read.csv(<dataDir>/<dataSet>/Courbe.csv)`, column 1 depth, column 2
  signal) and what synthData.R:16 writes (`write.csv(cbind(x,y), row.names=FALSE)`), for a whole directory tree.
* Stan-CSV writer / reader — the per-chain files `rstan::read_stan_csv` turns into a genuine `stanfit` (the route the R
  wrapper takes, r-pkg/R/fit.R); column order `lp__`, the six sampler parameters, then the model parameters with R's
  `name.index` flattening.
* `Results/<tag>_ctrl.txt` writer — the text block plotExpGP.R:4-23 sinks (`print(fit, pars=...)`).

Pure host-side formatting: no numerics of the path live here.
"""
from __future__ import annotations

import csv
import os

import numpy as np

from . import _abi as abi


def selX(x, y, depthSel=None, subSample=1):
    """FitOCTLib::selX as the reference calls it (FitOCT.R:85): optional depth window, then every subSample-th point."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    if depthSel is not None:
        keep = (x >= depthSel[0]) & (x <= depthSel[1])
        x, y = x[keep], y[keep]
    s = max(1, int(subSample))
    return x[::s], y[::s]


def read_courbe_csv(path: str):
    """One `Courbe.csv`: header line, then depth, signal (extra columns ignored)."""
    with open(path, newline="") as fh:
        rows = list(csv.reader(fh))
    if len(rows) < 2:
        raise ValueError(f"{path}: no data rows")
    body = rows[1:] if not _is_number(rows[0][0]) else rows
    a = np.array([[float(r[0]), float(r[1])] for r in body if len(r) >= 2 and r[0] != ""], dtype=np.float64)
    return a[:, 0], a[:, 1]


def _is_number(s: str) -> bool:
    try:
        float(s)
        return True
    except ValueError:
        return False


def read_data_dir(dataDir: str):
    """All `<dataDir>/<dataSet>/Courbe.csv` (FitOCT.R:74-84 loops over `list.dirs(dataDir)[-1]`), sorted by set name.
    Returns [(tag, x, y)], tag = '<dataDir>_<dataSet>' as FitOCT.R:80 builds it."""
    out = []
    base = os.path.basename(os.path.normpath(dataDir))
    for name in sorted(os.listdir(dataDir)):
        f = os.path.join(dataDir, name, "Courbe.csv")
        if os.path.isfile(f):
            x, y = read_courbe_csv(f)
            out.append((f"{base}_{name}", x, y))
    return out


def write_courbe_csv(path: str, x, y):
    """`write.csv(cbind(x,y), row.names=FALSE)` (synthData.R:16)."""
    with open(path, "w", newline="") as fh:
        w = csv.writer(fh, quoting=csv.QUOTE_NONNUMERIC)
        w.writerow(["x", "y"])
        for a, b in zip(x, y):
            w.writerow([float(a), float(b)])


def _flat_names(par_names):
    return [n.replace("[", ".").replace("]", "") for n in par_names]


def write_stan_csv(fit, prefix: str, model_name: str = "fitoct_b200"):
    """One Stan-CSV per chain: `<prefix>_<chain>.csv`.  `fit` is an api.StanFit with warm-up saved or not."""
    n_saved, chains, P = fit.draws.shape
    names = fit.par_names
    lp_col = names.index("lp__")
    other = [i for i in range(P) if i != lp_col]
    header = ["lp__"] + list(abi.SAMPLER_PARAM_NAMES) + _flat_names([names[i] for i in other])
    n_post = fit.n_iter - fit.n_warmup
    paths = []
    for c in range(chains):
        path = f"{prefix}_{c + 1}.csv"
        with open(path, "w") as fh:
            fh.write(f"# model = {model_name}\n# id = {c + 1}\n# method = sample (Default)\n")
            fh.write(f"#   sample\n#     num_samples = {n_post}\n#     num_warmup = {fit.n_warmup}\n")
            fh.write(f"#     save_warmup = {int(fit.save_warmup)}\n#     thin = 1\n")
            fh.write("#     algorithm = hmc (Default)\n#       engine = nuts (Default)\n#       metric = diag_e (Default)\n")
            fh.write(",".join(header) + "\n")
            for t in range(n_saved):
                row = [fit.draws[t, c, lp_col]] + list(fit.sampler_params[t, c]) + [fit.draws[t, c, i] for i in other]
                fh.write(",".join(repr(float(v)) if np.isfinite(v) else ("nan" if np.isnan(v) else ("inf" if v > 0 else "-inf"))
                                  for v in row) + "\n")
                if fit.save_warmup and t == fit.n_warmup - 1:
                    fh.write(_adaptation_block(fit, c))
            if not fit.save_warmup or fit.n_warmup == 0:
                fh.write(_adaptation_block(fit, c))
        paths.append(path)
    return paths


def _adaptation_block(fit, c):
    im = ", ".join(repr(float(v)) for v in fit.inv_metric[c])
    return (f"# Adaptation terminated\n# Step size = {float(fit.stepsize[c])!r}\n"
            f"# Diagonal elements of inverse mass matrix:\n# {im}\n")


def read_stan_csv(paths):
    """Minimal reader of the files written above (round-trip check; rstan's own reader is used on the R side).
    Returns dict(par_names, draws [n, chains, P], sampler_params [n, chains, 6], stepsize [chains], inv_metric)."""
    all_d, all_s, eps, invm, names = [], [], [], [], None
    for path in paths:
        rows, header = [], None
        with open(path) as fh:
            lines = fh.read().splitlines()
        for k, line in enumerate(lines):
            if line.startswith("# Step size"):
                eps.append(float(line.split("=")[1]))
            elif line.startswith("# Diagonal elements"):
                invm.append([float(v) for v in lines[k + 1].lstrip("# ").split(",")])
            if not line or line.startswith("#"):
                continue
            if header is None:
                header = line.split(",")
                continue
            rows.append([float(v) for v in line.split(",")])
        a = np.array(rows)
        names = header[7:] + ["lp__"]
        all_d.append(np.concatenate([a[:, 7:], a[:, :1]], axis=1))
        all_s.append(a[:, 1:7])
    return dict(par_names=names, draws=np.stack(all_d, axis=1), sampler_params=np.stack(all_s, axis=1),
                stepsize=np.array(eps), inv_metric=np.array(invm))


def write_ctrl_txt(path: str, fit, pars=("theta", "yGP", "lambda", "sigma", "br"), title="ExpGP parameters", append=True):
    """The block plotExpGP.R:4-23 sinks into `Results/<tag>_ctrl.txt`: a title line and rstan's summary table."""
    s = fit.summary(list(pars))
    with open(path, "a" if append else "w") as fh:
        fh.write(f"\n {title}:\n")
        fh.write(f"Inference for Stan model: fitoct_b200.\n{fit.draws.shape[1]} chains, each with iter={fit.n_iter}; "
                 f"warmup={fit.n_warmup}; thin=1;\n\n")
        cols = s["colnames"][:10]
        fh.write(f"{'':10s}" + "".join(f"{c:>10s}" for c in cols) + "\n")
        for name, row in zip(s["rownames"], s["summary"]):
            fh.write(f"{name:10s}" + "".join(f"{v:10.4g}" for v in row[:10]) + "\n")
        fh.write("\n\n")
