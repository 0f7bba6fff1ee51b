#!/usr/bin/env python
"""sass_sched.py — offline proxy for the issue behaviour of a SASS loop (no GPU needed).

Reads `cuobjdump -sass` output, finds the loops (backward branches) of a kernel, and runs a small in-order,
scoreboarded issue model of one SM sub-partition over a loop body: W warps execute the same body, one
instruction issues per cycle, an instruction waits for its source registers, the fp64 pipe accepts one warp
instruction every FP64_ISSUE cycles.  Latencies are the round-1 microbenchmark numbers
(profiles/r1_microbench.txt: dependent DFMA 8.4 cycles, 2.06-3.0 cycles issue).  It is a ranking tool for
schedule variants (is the sweep interleaved?  how long is the dependent chain?), not a predictor of absolute time.

  python scripts/sass_sched.py build/inst_10.o nuts_kernelILi10ELi0 [--warps 3] [--list]
"""
from __future__ import annotations

import argparse
import re
import subprocess
import sys
from collections import Counter

LINE = re.compile(r"^\s+/\*([0-9a-f]+)\*/\s+(.*?);")
REG = re.compile(r"\bR(\d+)\b")
PRED = re.compile(r"\b(U?P\d)\b")

FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")
LAT = {"fp64": 8.4, "alu": 4.5, "lds": 29.0, "ldg": 45.0, "mufu": 20.0, "shfl": 25.0, "ldl": 40.0}


def disasm(obj, fun):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    blocks, cur, name = {}, None, None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            name = m.group(1)
            cur = blocks.setdefault(name, [])
            continue
        m = LINE.match(ln)
        if m and cur is not None:
            cur.append((int(m.group(1), 16), m.group(2).strip()))
    hits = [k for k in blocks if fun in k]
    if not hits:
        sys.exit(f"no function matching {fun}; have: {list(blocks)[:8]}")
    return hits[0], blocks[hits[0]]


def parse(text):
    """-> (opcode, klass, dst regs, src regs, dst preds, src preds)"""
    guard = None
    t = text
    m = re.match(r"@(!?U?P\d)\s+(.*)", t)
    if m:
        guard, t = m.group(1).lstrip("!"), m.group(2)
    op, _, rest = t.partition(" ")
    base = op.split(".")[0]
    ops = [o.strip() for o in rest.split(",")] if rest else []
    wide = base in FP64 or ".64" in op or base == "CS2R" or "WIDE" in op
    wide128 = ".128" in op

    def regs_of(o, w):
        rs = []
        for r in REG.findall(o):
            r = int(r)
            n = 4 if (w and wide128) else (2 if w else 1)
            rs += list(range(r, r + n))
        return rs

    if base in FP64:
        klass = "fp64"
    elif base == "LDS":
        klass = "lds"
    elif base in ("LDG", "LD"):
        klass = "ldg"
    elif base in ("LDL", "STL"):
        klass = "ldl"
    elif base == "MUFU":
        klass = "mufu"
    elif base == "SHFL":
        klass = "shfl"
    else:
        klass = "alu"
    dst, src, dp, sp = [], [], [], []
    store = base in ("STS", "STG", "STL", "ST", "BRA", "BSSY", "BSYNC", "ISETP", "DSETP", "FSETP", "BAR", "EXIT")
    for i, o in enumerate(ops):
        is_dst = i == 0 and not store
        if base in ("DSETP", "ISETP", "FSETP") and i < 2:
            dp += PRED.findall(o)
            continue
        if base == "IADD3" and i in (1, 2) and PRED.fullmatch(o or ""):
            dp += [o]
            continue
        w = wide and not (base == "MUFU") and not (base in ("LDS", "LDG", "LDL") and i > 0)
        if base in ("LDS", "LDG", "LDL") and i > 0:
            w = ".64" in o  # address pair only for 64-bit addressing
        if is_dst:
            dst += regs_of(o, w if base != "MUFU" else False)
        else:
            src += regs_of(o, w)
            sp += [p for p in PRED.findall(o) if p not in ("PT",)]
    if guard:
        sp.append(guard)
    # register-file bandwidth (profiles/r1_microbench.txt, issue_cost.cu): an fp64 instruction with three distinct
    # register-pair sources occupies the pipe 3.0 cycles, 2.38 when one of them comes from the operand-reuse cache
    cost = 2.06
    if klass == "fp64":
        pairs = [o for o in ops[1:] if REG.search(o)]
        if len({REG.search(o).group(1) for o in pairs}) >= 3:
            cost = 3.0
    return op, klass, dst, src, dp, sp, cost, [REG.search(o).group(1) for o in ops[1:] if ".reuse" in o and REG.search(o)]


def loops(ins):
    addr2i = {a: i for i, (a, _) in enumerate(ins)}
    out = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA(?:\.\S+)?\s+(?:\S+,\s*)?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr2i:
                out.append((addr2i[tgt], i))
    return out


def simulate(body, warps, fp64_issue=2.06, iters=6, bank_model=True):
    parsed = [parse(t) for _, t in body]
    n = len(parsed)
    ready = [dict() for _ in range(warps)]  # reg/pred -> cycle its value is available
    pc = [0] * warps
    it = [0] * warps
    t_done = [0.0] * warps
    t_start_last = [0.0] * warps
    fp64_free = 0.0
    cyc = 0.0
    last = 0
    while min(it) < iters:
        issued = False
        for k in range(warps):
            w = (last + 1 + k) % warps
            if it[w] >= iters:
                continue
            op, klass, dst, src, dp, sp, cost, reuse = parsed[pc[w]]
            need = max([ready[w].get(("r", r), 0.0) for r in src] + [ready[w].get(("p", p), 0.0) for p in sp] + [0.0])
            # WAW/WAR on in-flight loads ignored
            if need > cyc:
                continue
            if klass == "fp64" and fp64_free > cyc + 0.99:
                continue
            if klass == "fp64":
                c = cost
                if c == 3.0 and pc[w] > 0 and any(str(r // 1) in map(str, parsed[pc[w] - 1][7]) for r in src):
                    c = 2.38
                fp64_free = max(fp64_free, cyc) + (c if bank_model else fp64_issue)
            lat = LAT[klass]
            for r in dst:
                ready[w][("r", r)] = cyc + lat
            for p in dp:
                ready[w][("p", p)] = cyc + lat
            pc[w] += 1
            if pc[w] == n:
                pc[w] = 0
                it[w] += 1
                if it[w] == iters - 1:
                    t_start_last[w] = cyc
                if it[w] == iters:
                    t_done[w] = cyc
            last = w
            issued = True
            break
        cyc += 1.0
        if not issued and cyc > 1e7:
            break
    per_iter = max(t_done) / iters
    return per_iter


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("obj")
    ap.add_argument("fun")
    ap.add_argument("--warps", type=int, nargs="*", default=[1, 3, 4])
    ap.add_argument("--list", action="store_true", help="print the body of the selected loop")
    ap.add_argument("--min-fp64", type=int, default=30, help="only loops with at least this many fp64 instructions")
    ap.add_argument("--all", action="store_true")
    a = ap.parse_args()
    name, ins = disasm(a.obj, a.fun)
    print(f"# {name}: {len(ins)} instructions")
    seen = set()
    for lo, hi in sorted(loops(ins), key=lambda x: x[0]):
        body = ins[lo:hi + 1]
        ops = Counter(parse(t)[0].split(".")[0] for _, t in body)
        nf = sum(ops[o] for o in FP64)
        if nf < a.min_fp64 or any("BAR" in t or "CALL" in t for _, t in body):
            continue
        if len(body) > 1200 and not a.all:
            continue
        key = (len(body), nf)
        if key in seen and not a.all:
            continue
        seen.add(key)
        sims = {w: simulate(body, w) for w in a.warps}
        mufu = [i for i, (_, t) in enumerate(body) if "MUFU.RCP64H" in t]
        pp = [parse(t) for _, t in body]
        pipe = 0.0
        for i, q in enumerate(pp):
            if q[1] == "fp64":
                c = q[6]
                if c == 3.0 and i > 0 and any(str(r) in pp[i - 1][7] for r in q[3]):
                    c = 2.38
                pipe += c
        print(f"loop 0x{ins[lo][0]:x}-0x{ins[hi][0]:x}: {len(body)} instr, fp64 {nf} "
              f"(DFMA {ops['DFMA']} DMUL {ops['DMUL']} DADD {ops['DADD']} DSETP {ops['DSETP']}), LDS {ops['LDS']}, "
              f"FSEL {ops['FSEL']}, CS2R {ops['CS2R']}, LDL/STL {ops['LDL']}/{ops['STL']}; rcp at {mufu}; fp64 pipe {pipe:.0f} cyc; "
              + ", ".join(f"{w}w: {c:.0f} cyc/iter ({len(body) * w / c:.2f} ipc)" for w, c in sims.items()))
        if a.list:
            for ad, t in body:
                print(f"   {ad:05x} {t}")


if __name__ == "__main__":
    main()
