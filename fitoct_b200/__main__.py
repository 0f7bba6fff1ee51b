"""Batch driver: what FitOCT.R:72-131 does for every `<dataDir>/<dataSet>/Courbe.csv`, with the whole directory handed to
the GPU library at once.

    python -m fitoct_b200 <dataDir> [--ctrl ctrlParams.yaml] [--out Results] [--chains 4] [--seed 1234] [--no-gate]

Per data set it writes `<out>/<tag>_ctrl.txt` (the text FitOCT.R sinks through plotMonoExp.R:2-11 and plotExpGP.R:4-23)
and, for method 'sample', one Stan-CSV per chain (`<out>/<tag>_ExpGP_<chain>.csv`, readable by rstan::read_stan_csv).
Plots are the reference's business (plot*.R) and are not produced.  NB: FitOCT.R `break`s out of its loop at the first
data set whose mono-exponential fit passes the Birge-ratio test (FitOCT.R:100); this driver goes on to the next one.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

from . import _abi as abi
from . import api
from . import io as fio


def _mono_block(theta, br, ci, alert):
    lines = ["", " MonoExp decay parameters:"]
    lines += [f"b_{i + 1} :  {theta[i]:.3g} " for i in range(3)]            # plotMonoExp.R:4-7
    lines += ["", "", f"br   : {br:.3g}", f"CI95 : [{ci[0]:.3g}, {ci[1]:.3g}]"]
    if alert:
        lines.append("!!! WARNING !!! br out of interval")
    return "\n".join(lines) + "\n"


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m fitoct_b200", description=__doc__.split("\n\n")[0])
    ap.add_argument("dataDir")
    ap.add_argument("--ctrl", default="ctrlParams.yaml", help="YAML overrides of the FitOCT.R:37-53 defaults")
    ap.add_argument("--out", default="Results")
    ap.add_argument("--chains", type=int, default=4)
    ap.add_argument("--seed", type=int, default=1234)     # FitOCT.R:15
    ap.add_argument("--no-gate", action="store_true", help="fit the modulated model to every data set")
    a = ap.parse_args(argv)
    ctrl = api.load_ctrl_params(a.ctrl)
    print("Configuration Parameters\n------------------------")       # FitOCT.R:64-66
    for k, v in ctrl.items():
        print(f" $ {k:12s}: {v}")
    sets = fio.read_data_dir(a.dataDir)
    if not sets:
        print(f"no <dataSet>/Courbe.csv under {a.dataDir}", file=sys.stderr)
        return 1
    os.makedirs(a.out, exist_ok=True)
    # one library call per depth grid (profiles of a directory normally share it)
    groups: dict = {}
    for tag, x, y in sets:
        groups.setdefault(x.tobytes(), []).append((tag, x, y))
    Nn = int(ctrl["Nn"])
    for members in groups.values():
        x = members[0][1]
        Y = np.stack([m[2] for m in members])
        out = api.FitOCT_batch(x, Y, ctrl, chains=a.chains, seed=a.seed, gate=not a.no_gate, keep_draws=True)
        where = {int(j): k for k, j in enumerate(out["expgp_index"])}
        for j, (tag, _, _) in enumerate(members):
            path = os.path.join(a.out, f"{tag}_ctrl.txt")
            alert = bool(out["alert"][j])
            with open(path, "a") as fh:
                fh.write(_mono_block(out["mono_theta"][j], out["mono_br"][j], out["br_ci"][j], alert))
            status = "MonoExp fit OK"
            if j in where:
                k = where[j]
                e = out["expgp"]
                if out["method"] == "sample":
                    fit = api._stanfit_from(e, k, abi.FOCT_EXPGP, Nn, out["sampler_cfg"])
                    fio.write_ctrl_txt(path, fit)
                    fio.write_stan_csv(fit, os.path.join(a.out, f"{tag}_ExpGP"))
                    br_gp = float(fit.summary_table[Nn + 5, 0])
                else:
                    row = e["par"][k] if out["method"] == "optim" else e["mean"][k]
                    with open(path, "a") as fh:
                        fh.write("\n ExpGP parameters:\n")                    # plotExpGP.R:13-17
                        fh.write(f"theta  : {' '.join(f'{v:.6g}' for v in row[:3])} \n")
                        fh.write(f"yGP    : {' '.join(f'{v:.6g}' for v in row[3:3 + Nn])} \n")
                        fh.write(f"lambda : {row[3 + Nn]:.6g} \nsigma  : {row[4 + Nn]:.6g} \nbr     : {row[5 + Nn]:.6g} \n\n\n")
                    br_gp = float(row[5 + Nn])
                status = f"ExpGP ({out['method']}) br = {br_gp:.3g}"
            print(f"{tag} ------------------------ MonoExp br = {out['mono_br'][j]:.3g}"
                  f"{' (alert)' if alert else ''}; {status}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
