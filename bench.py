#!/usr/bin/env python
"""bench.py — fitExpGP batch throughput (BASELINE.json metric: min-ESS/s and draws/s) on B200.

One "step" = one complete NUTS fit (warm-up + sampling + on-device summary) of a batch of synthetic
synthData.R-shaped profiles, 4 chains each (BASELINE.json configs[2], `config.workload`).  Weak scaling: every
rank (one process per GPU, launched by torchrun for N > 1) fits its own shard of `--profiles` profiles; the
shards are independent (no data-path collective) — torch.distributed is used only for the
barrier and the max-over-ranks timing.

  value      device-resident inputs (foct_plan_*), timed with CUDA events on the plan's stream
  e2e        the one-shot C-ABI call foct_sample() with HOST buffers, H2D/D2H inside the timed region
  roofline   fp64 FMA roofline of the sampling kernel: counted leapfrogs x F_grad(N, Nn) / kernel time
             vs the DFMA peak measured in this run (MEASURED_PEAKS.json has no fp64 figure)
  cpu_baseline / --impl reference
             the CPU oracle (Stan's algorithm, analytic gradient — NOT rstan, which cannot run here:
             BASELINE.md §3) on a bounded sample of the same workload, all host cores; its TIMING build
             (-O3, AVX2 + FMA: oracle/Makefile), not the separately-rounded parity build.
  value/e2e  the fixed-length workload BASELINE configs[2] names (500 + 1000 iterations), the same work the CPU arm does.
  until_converged
             north_star's "sampled to R-hat < 1.01": the same fit with rhat_target = 1.01 — the profiles still above it
             after the configured iterations are continued on the device (rhat_target / max_extend / extend_iter) and
             re-summarised; one extra timed step, reported next to the headline with its own R-hat figures.
  inlib      (N > 1) the product's own multi-GPU path: rank 0 alone hands all N x profiles to ONE foct_sample()
             call with devices = 0..N-1 (one host thread per GPU inside the library) while the other ranks wait.
  c5         (N = 8, or --c5) BASELINE configs[4]: 100,000 profiles x 4 chains through that same call.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from fitoct_b200 import _abi as abi  # noqa: E402
from fitoct_b200 import synth  # noqa: E402

METRIC = "fitExpGP batch min-ESS/s"
UNIT = "ESS/s"


def f_grad(N: int, Nn: int) -> float:
    """Algorithmic flops of one gradient evaluation, SURVEY §8(d): N (4 Nn + 23)."""
    return float(N) * (4.0 * Nn + 23.0)


def make_batch(n, first_id, Nn):
    S = synth.make_profiles(n, first_id=first_id, modulated_only=True)
    b = abi.make_problems_dense(S["x"], S["Y"], S["UY"], S["theta0"], S["Sigma0"], Nn=Nn, ids=S["ids"])
    return S, b


def measured_traffic(n, Nn, n_warmup, n_iter, chains):
    """DRAM bytes of one sampling-kernel launch from the committed ncu capture, if it was taken on this workload."""
    here = os.path.dirname(os.path.abspath(__file__))
    for name in ("r2_dram_traffic.json", "r1_dram_traffic.json"):
        try:
            with open(os.path.join(here, "profiles", name)) as fh:
                t = json.load(fh)
        except (OSError, ValueError):
            continue
        w = t.get("workload", {})
        if (w.get("profiles_per_gpu"), w.get("Nn"), w.get("n_warmup"), w.get("n_iter"), w.get("chains")) != (n, Nn, n_warmup, n_iter, chains):
            return None, f"ncu capture is for another workload (profiles/{name})"
        return t["dram_bytes_per_launch"], f"profiles/{name} (ncu dram__bytes_read.sum + dram__bytes_write.sum, one launch of {t.get('kernel')})"
    return None, "no ncu capture committed"


def min_ess_sum(summary, Nn):
    """Sum over profiles of min over the Nn+5 sampled parameters of Bulk_ESS."""
    cols = list(range(0, Nn + 5))  # theta, yGP, lambda, sigma  (br, lp__ are derived)
    be = summary[:, cols, 10]
    return float(np.nansum(np.nanmin(be, axis=1)))


def host_cores() -> int:
    """Cores this process may really use: affinity mask capped by the cgroup CPU quota (torchrun also exports
    OMP_NUM_THREADS=1, so the thread count is always passed to the oracle explicitly)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        with open("/sys/fs/cgroup/cpu.max") as fh:
            quota, period = fh.read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(float(quota) / float(period) + 0.5)))
    except Exception:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            p_ = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, int(q / p_ + 0.5)))
        except Exception:
            pass
    return max(1, n)


class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def start(self):
        def run():
            while not self._stop.is_set():
                try:
                    o = subprocess.run(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                        "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                    self.rows.append([c.strip() for c in o.strip().split(",")])
                except Exception:
                    pass
                self._stop.wait(0.5)
        self._t = threading.Thread(target=run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for k, nme in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args, rank, world):
    """--impl reference: the CPU restatement on host cores (rank 0 only)."""
    if rank != 0:
        return
    from oracle import oracle as O
    cores = host_cores()
    chains = 4
    n = max(1, min(args.profiles, max(1, cores // chains) * args.cpu_waves))
    _, b = make_batch(n, 0, args.nn)
    spec = abi.default_spec()
    cfg = abi.default_cfg(n_warmup=args.n_warmup, n_iter=args.n_iter, seed=args.seed, chains=chains)
    times, ess, threads = [], 0.0, cores
    for s in range(args.warmup + args.steps):
        cfg.seed = args.seed + s
        t0 = time.perf_counter()
        out = O.sample(abi.FOCT_EXPGP, b, n, spec, cfg, draws=True, summary=True, n_threads=cores, fast=True)
        dt = time.perf_counter() - t0
        threads = out["threads"]
        if s >= args.warmup:
            times.append(dt)
            ess = min_ess_sum(out["summary"], args.nn)
    T = float(np.sum(times))
    n_post = args.n_iter - args.n_warmup
    val = ess * len(times) / T
    dps = n * chains * n_post * len(times) / T
    sample = (f"{n} of {args.profiles} profiles x {chains} chains, full {args.n_warmup}/{n_post} iterations per step, fixed length "
              f"(no run-until-converged rounds)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * T / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic (synthData.R-shaped)", "draws_per_s": dps,
        "config": workload_config(args, args.profiles),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "draws_per_s": dps,
                         "build": "oracle/libfoct_oracle_fast.so: gcc -O3 -march=x86-64-v3 (AVX2 + FMA), OpenMP over chains",
                         "note": "CPU restatement (Stan algorithm, analytic gradient) - not rstan (R absent, BASELINE.md s3)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n):
    return {"workload": f"batch of {n} synthetic synthData.R profiles x 4 chains, fitExpGP Nn={args.nn}, "
                        f"warmup {args.n_warmup} + {args.n_iter - args.n_warmup} draws (BASELINE configs[2])",
            "profiles_per_gpu": n, "chains": 4, "Nn": args.nn, "N": 481, "n_warmup": args.n_warmup,
            "n_iter": args.n_iter, "adapt_delta": 0.8, "max_treedepth": 10,
            "l2": "flushed between steps (256 MiB memset)", "parallelism": "independent profile shards, no collective"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--profiles", type=int, default=1000, help="profiles per GPU")
    ap.add_argument("--nn", type=int, default=10)
    ap.add_argument("--n-warmup", type=int, default=500)
    ap.add_argument("--n-iter", type=int, default=1500)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--cpu-waves", type=int, default=3, help="CPU sample = cores/4 * waves profiles")
    ap.add_argument("--rhat-target", type=float, default=1.01, help="continue profiles until split R-hat < this (0: off)")
    ap.add_argument("--max-extend", type=int, default=12, help="continuation rounds of (n_iter - n_warmup) / 4 draws")
    ap.add_argument("--c5", action="store_true", help="also run BASELINE configs[4] (100k profiles) on all N GPUs from rank 0")
    ap.add_argument("--c5-profiles", type=int, default=100000)
    ap.add_argument("--no-inlib", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from fitoct_b200 import _lib as L

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: fitoct_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo", rank=rank, world_size=world)  # barrier/timing only; shards never talk

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from fitoct_b200 import shard

    chains = 4
    n = args.profiles
    n_post = args.n_iter - args.n_warmup
    S, batch = make_batch(n, rank * n, args.nn)
    spec = abi.default_spec()
    import ctypes as C
    dev_arr = (C.c_int * 1)(local_rank)

    def make_cfg(extend=False):
        c = abi.default_cfg(n_warmup=args.n_warmup, n_iter=args.n_iter, seed=args.seed, chains=chains)
        c.n_devices = 1
        c.devices = C.cast(dev_arr, C.POINTER(C.c_int))
        if extend and args.rhat_target > 0:
            c.rhat_target, c.max_extend = args.rhat_target, args.max_extend   # rounds of n_post / 4 draws (extend_iter = 0)
        return c

    cfg = make_cfg()
    dist_or_none = dist if world > 1 else None
    peak_tf, _ = L.fp64_peak(local_rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timed_plan(cfg_, steps, warmup, sample_clocks):
        """`steps` timed fits with device-resident inputs; CUDA-event times of the kernels on the plan's stream."""
        plan = L.Plan(abi.FOCT_EXPGP, batch, n, spec, cfg_, want_draws=False, want_summary=True)
        for s in range(warmup):
            flush.zero_()
            plan.run(args.seed + s)
            plan.sync()
        clocks = ClockSampler(local_rank) if sample_clocks else None
        barrier()
        if clocks:
            clocks.start()
        r = dict(step_ms=[], samp_ms=[], summ_ms=[], ess=0.0, leap=0.0, launches=0, extended=0.0)
        t0 = time.perf_counter()
        for s in range(steps):
            flush.zero_()
            torch.cuda.synchronize()
            plan.run(args.seed + args.warmup + s)
            plan.sync()
            tm = plan.timing()
            r["samp_ms"].append(tm["sample_ms"]); r["summ_ms"].append(tm["summary_ms"])
            r["step_ms"].append(tm["sample_ms"] + tm["summary_ms"])
            out = plan.fetch()
            r["ess"] += min_ess_sum(out["summary"], args.nn)
            r["leap"] += float(out["n_leapfrog"].sum())
            r["launches"] += plan.launches()
            r["extended"] += float((out["n_extend"] > 0).sum())
        barrier()
        r["wall"] = time.perf_counter() - t0
        r["clk"] = clocks.stop() if clocks else None
        r["tm"], r["out"] = plan.timing(), out
        plan.close()
        return r

    # ---------------- value: device-resident inputs, CUDA-event timing
    R = timed_plan(cfg, args.steps, args.warmup, True)
    out, tm, clk, t_wall = R["out"], R["tm"], R["clk"], R["wall"]
    rh = np.nanmax(out["summary"][:, : args.nn + 5, 9], axis=1)
    rhat_max, rhat_q99 = float(np.nanmax(rh)), float(np.nanquantile(rh, 0.99))
    n_div = float(out["n_divergent"].sum())
    T = float(np.sum(R["step_ms"])) * 1e-3  # device time of the timed steps on this rank
    mx, sm = shard.aggregate([T, float(np.sum(R["samp_ms"])) * 1e-3], [R["ess"], R["leap"], R["extended"], float((rh < 1.01).sum())],
                             dist_or_none)
    T_max, Ts_max = float(mx[0]), float(mx[1])
    ess_all, leap_all, ext_all, conv_all = (float(v) for v in sm)
    value = ess_all / T_max
    draws_per_s = world * n * chains * n_post * args.steps / T_max

    # ---------------- the same fit run until every profile is below the R-hat target (north_star), one timed step
    conv = None
    if args.rhat_target > 0:
        F = timed_plan(make_cfg(True), 1, 1, False)
        rhf = np.nanmax(F["out"]["summary"][:, : args.nn + 5, 9], axis=1)
        mxf, smf = shard.aggregate([float(np.sum(F["step_ms"])) * 1e-3, float(np.nanmax(rhf))],
                                   [F["ess"], float((rhf < args.rhat_target).sum()), F["extended"], F["leap"]], dist_or_none)
        conv = {"value": float(smf[0]) / float(mxf[0]), "unit": UNIT, "ms_per_step": 1e3 * float(mxf[0]),
                "rhat_target": args.rhat_target, "max_extend": args.max_extend, "rhat_max": float(mxf[1]),
                "profiles_below_target": float(smf[1]) / (world * n), "profiles_continued": float(smf[2]) / world,
                "grad_per_s": float(smf[3]) / float(mxf[0]), "gpu_launches": int(F["launches"])}

    # ---------------- e2e: host buffers through foct_sample (H2D + D2H inside the timed region)
    e2e = e2e_draws = single = None
    if not args.no_e2e:
        D, P_out = abi.dims(abi.FOCT_EXPGP, args.nn)
        h2d = n * (3 * 481 * 8 + 200)
        d2h = n * (P_out * abi.FOCT_N_SUMMARY_COLS * 8 + chains * 8 * (1 + D + 2 + 1 + D) + 4)

        def timed_e2e(steps, draws):
            L.sample(abi.FOCT_EXPGP, batch, min(n, 64), spec, cfg, draws=draws, summary=True)  # warm the path
            barrier()
            t0 = time.perf_counter()
            e_ess = 0.0
            for s in range(steps):
                cfg.seed = args.seed + args.warmup + s
                o2 = L.sample(abi.FOCT_EXPGP, batch, n, spec, cfg, draws=draws, summary=True)
                e_ess += min_ess_sum(o2["summary"], args.nn)
            barrier()
            Te = time.perf_counter() - t0
            cfg.seed = args.seed
            m2, s2 = shard.aggregate([Te], [e_ess], dist_or_none)
            return float(m2[0]), float(s2[0])

        Te, e_all = timed_e2e(args.steps, False)
        e2e = {"value": e_all / Te, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "draws_per_s": world * n * chains * n_post * args.steps / Te, "ms_per_step": 1e3 * Te / args.steps}
        # the reference contract returns the draws (plotExpGP.R:44-47): the same call shipping them and the sampler
        # parameters to the host as well
        Td, d_all = timed_e2e(1, True)
        # BASELINE configs[1]: ONE profile x 4 chains through the same call with host buffers - what FitOCT.R's loop
        # submits per Courbe.csv (FitOCT.R:110-124); runs on the latency kernel (two warps per chain)
        if rank == 0:
            one = abi.make_problems_dense(S["x"], S["Y"][:1], S["UY"][:1], S["theta0"][:1], S["Sigma0"][:1], Nn=args.nn, ids=S["ids"][:1])
            L.sample(abi.FOCT_EXPGP, one, 1, spec, cfg, draws=True, summary=True)
            t0 = time.perf_counter()
            o1 = L.sample(abi.FOCT_EXPGP, one, 1, spec, cfg, draws=True, summary=True)
            t1 = time.perf_counter() - t0
            single = {"wall_s": t1, "leapfrogs": float(o1["n_leapfrog"].sum()), "min_bulk_ess": float(np.nanmin(o1["summary"][0, : args.nn + 5, 10])),
                      "rhat_max": float(np.nanmax(o1["summary"][0, : args.nn + 5, 9])), "kernel": "nuts_lat_kernel (two warps per chain)"}
        e2e_draws = {"value": d_all / Td, "unit": UNIT, "ms_per_step": 1e3 * Td, "h2d_bytes_per_step": h2d,
                     "d2h_bytes_per_step": d2h + n * chains * n_post * (P_out + 6) * 8}

    # ---------------- the product's own multi-GPU path: one foct_sample(devices = 0..N-1) issued by rank 0
    inlib = c5 = None
    run_c5 = args.c5 or world == 8
    if world > 1 and (not args.no_inlib or run_c5):
        barrier()
        if rank == 0:
            devs = list(range(world))
            if not args.no_inlib:
                _, ball = make_batch(world * n, 0, args.nn)
                ci = make_cfg()
                # warm the path with the SAME call: a smaller batch would run on the other sampling kernel (chosen by batch
                # size) and leave the module load of this one on devices 1..N-1 inside the timed call (0.7 s at N = 2)
                L.sample(abi.FOCT_EXPGP, ball, world * n, spec, ci, draws=False, summary=True, devices=devs)
                t0 = time.perf_counter()
                oi = L.sample(abi.FOCT_EXPGP, ball, world * n, spec, ci, draws=False, summary=True, devices=devs)
                Ti = time.perf_counter() - t0
                inlib = {"value": min_ess_sum(oi["summary"], args.nn) / Ti, "unit": UNIT, "ms_per_step": 1e3 * Ti, "n_gpus": world,
                         "profiles": world * n, "how": "ONE foct_sample() call from rank 0 with devices=0..N-1, host buffers, wall clock",
                         "rhat_max": float(np.nanmax(oi["summary"][:, : args.nn + 5, 9]))}
                del ball, oi
            if run_c5:
                n5 = args.c5_profiles
                _, b5 = make_batch(n5, 0, args.nn)
                c5cfg = make_cfg(True)   # north_star: sampled to R-hat < 1.01
                t0 = time.perf_counter()
                o5 = L.sample(abi.FOCT_EXPGP, b5, n5, spec, c5cfg, draws=False, summary=True, devices=devs)
                T5 = time.perf_counter() - t0
                rh5 = np.nanmax(o5["summary"][:, : args.nn + 5, 9], axis=1)
                c5 = {"workload": f"{n5} profiles x {chains} chains, Nn={args.nn}, {args.n_warmup}+{n_post} iterations, ONE foct_sample() "
                                  f"call over {world} GPUs (BASELINE configs[4], strong scaling)",
                      "seconds": T5, "value": min_ess_sum(o5["summary"], args.nn) / T5, "unit": UNIT,
                      "draws_per_s": n5 * chains * n_post / T5, "grad_per_s": float(o5["n_leapfrog"].sum()) / T5,
                      "rhat_max": float(np.nanmax(rh5)), "profiles_below_1.01": float((rh5 < 1.01).mean()),
                      "profiles_continued": float((o5["n_extend"] > 0).mean()), "divergent_post_warmup": float(o5["n_divergent"].sum())}
                del b5, o5
        barrier()

    # ---------------- cpu baseline (rank 0, N = 1 only): bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        cores = host_cores()
        nc = max(1, min(n, max(1, cores // chains) * args.cpu_waves))
        _, bc = make_batch(nc, 0, args.nn)
        cpu_cfg = abi.default_cfg(n_warmup=args.n_warmup, n_iter=args.n_iter, seed=args.seed, chains=chains)
        t0 = time.perf_counter()
        oc = O.sample(abi.FOCT_EXPGP, bc, nc, spec, cpu_cfg, draws=True, summary=True, n_threads=cores, fast=True)
        dt = time.perf_counter() - t0
        cpu = {"value": min_ess_sum(oc["summary"], args.nn) / dt, "unit": UNIT, "cores": oc["threads"], "kind": "port",
               "sample": f"{nc} of {n} profiles x {chains} chains, full {args.n_warmup}/{n_post} iterations, {dt:.1f} s",
               "draws_per_s": nc * chains * n_post / dt, "grad_per_s": float(oc["n_leapfrog"].sum()) / dt,
               "build": "oracle/libfoct_oracle_fast.so: gcc -O3 -march=x86-64-v3 (AVX2 + FMA), OpenMP over chains",
               "note": "CPU restatement (Stan algorithm, analytic gradient) - not rstan (R absent, BASELINE.md s3)"}

    if rank == 0:
        achieved = leap_all / world * f_grad(481, args.nn) / Ts_max / 1e12  # per-GPU TFLOP/s of the sampling kernel
        traffic, traffic_src = measured_traffic(n, args.nn, args.n_warmup, args.n_iter, chains)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * T_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic (synthData.R-shaped, fitoct_b200/synth.py)",
            "draws_per_s": draws_per_s, "grad_per_s": leap_all / Ts_max,
            "config": workload_config(args, n),
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved / peak_tf, "traffic": traffic, "traffic_unit": "DRAM bytes per launch",
                         "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": n * ((3 + args.nn) * 512 * 8 + chains * n_post * (args.nn + 7) * 8),
                         "peak_source": "DFMA-chain microbenchmark measured in this run (foct_fp64_peak); "
                                        "MEASURED_PEAKS.json has no fp64 figure",
                         "kernel": ("foct::nuts2w_kernel (two chains per warp, warps claim (profile, chain pair) units, time-sliced)"
                                    if tm["block"] == 64 and args.nn <= 11 else "foct::nuts_kernel (one chain per warp)"),
                         "algorithmic_flop_per_grad": f_grad(481, args.nn),
                         "leapfrogs_per_step": leap_all / world / args.steps,
                         "hbm_writeback_gbs": (n * chains * n_post * (args.nn + 7) * 8 * args.steps / Ts_max) / 1e9,
                         "launch": {k: tm[k] for k in ("grid", "block", "blocks_per_sm", "regs", "smem_bytes")}},
            "kernel_ms": {"sample": float(np.mean(R["samp_ms"])), "summary": float(np.mean(R["summ_ms"]))},
            "quality": {"rhat_max": rhat_max, "rhat_q99_of_profile_max": rhat_q99, "divergent_post_warmup": n_div,
                        "profiles_below_1.01": conv_all / (world * n),
                        "profiles_continued_per_step": ext_all / (world * args.steps),
                        "mean_min_bulk_ess_per_profile": ess_all / (world * n * args.steps)},
            "until_converged": conv,
            "clocks": clk, "e2e": e2e, "e2e_with_draws": e2e_draws, "single_profile": single, "gpu_launches": int(R["launches"]),
            "wall_s_timed_region": t_wall,
        }
        if inlib is not None:
            line["inlib"] = inlib
        if c5 is not None:
            line["c5"] = c5
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
