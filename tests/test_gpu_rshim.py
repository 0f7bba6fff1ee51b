"""GPU (-m gpu): the R `.Call` shim (r-pkg/src/shim.c) compiled against the stand-in R runtime of tests/r_stub and RUN by
a C driver through the C ABI — R itself is absent from the image.  Covers SURVEY §8(b): one `.Call` for a
batch, rstan-style progress lines that the Shiny scraper can parse (ShinyInterface/server.R:457-472), an interrupt raised
between polls that cancels the kernels and comes back as an R error, and the library staying usable afterwards."""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_driver(tmp_path):
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    exe = str(tmp_path / "drive_shim")
    srcs = [os.path.join(ROOT, "r-pkg", "src", "shim.c"), os.path.join(ROOT, "tests", "r_stub", "rstub.c"),
            os.path.join(ROOT, "tests", "r_stub", "drive_shim.c")]
    libdir = os.path.join(ROOT, "fitoct_b200")
    r = subprocess.run([cc, "-std=gnu99", "-O1", "-I", os.path.join(ROOT, "tests", "r_stub"), "-I", os.path.join(ROOT, "include"),
                        *srcs, "-o", exe, "-L", libdir, "-lfitoct_b200", "-Wl,-rpath," + libdir, "-lm"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_shim_batch_progress_interrupt_pipeline(tmp_path, L):
    exe = build_driver(tmp_path)
    r = subprocess.run([exe, "3"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = r.stdout
    assert "DONE" in out and "UNEXPECTED_ERROR" not in out
    # 1. one .Call for the batch; draws sized [n, n_saved, chains, P_out]
    m = re.search(r"BATCH draws_len (\d+) expected (\d+) interrupt_checks (\d+)", out)
    assert m and m.group(1) == m.group(2) and int(m.group(3)) >= 1
    means = re.findall(r"BATCH_MEAN \d+ (\S+) (\S+) (\S+) rhat (\S+)", out)
    assert len(means) == 3
    for a, b, c, rh in means:   # synthData.R truth (1000, 2000, 300), 5 % priors
        assert abs(float(a) - 1000) < 30 and abs(float(b) - 2000) < 120 and abs(float(c) - 300) < 30 and float(rh) < 1.2
    # 2. progress lines in rstan's format; the Shiny scraper's arithmetic on the last line gives ~100 %
    lines = re.findall(r"Chain (\d+): Iteration:\s+(\d+) / (\d+) \[\s*(\d+)%\]\s+\((Warmup|Sampling|Extending)\)", out.split("BATCH draws_len")[0])
    assert len(lines) >= 2
    chain, it, n_iter, pct, _ = lines[-1]
    assert int(n_iter) == 200 and ((int(chain) - 1) * 100 + int(pct)) / 4 >= 99
    shown = [((int(c) - 1) * 100 + int(p)) / 4 for c, _, _, p, _ in lines]
    assert shown == sorted(shown)                     # monotone progress
    assert any(ph == "Warmup" for *_, ph in lines[:1]) or shown[0] > 50
    # 3. user interrupt at the 3rd poll: cancelled on the device, reported through Rf_error, nothing leaked into a hang
    assert re.search(r"INTERRUPT raised after 3 checks: fitoct_b200: interrupted by the user", out)
    # 4. library usable afterwards; single-profile entry == profile 0 of the batch; generated quantities consistent
    assert "SINGLE equals_batch_profile0 1" in out
    m = re.search(r"PREDICT n (\d+) max\|y-m-resid\| (\S+)", out)
    assert m and int(m.group(1)) == 2 * 481 and float(m.group(2)) < 1e-9
    m = re.search(r"SUMMARY rows (\d+) max_rel_diff (\S+)", out)
    assert m and int(m.group(1)) == 17 and float(m.group(2)) == 0.0   # foct_R_summary of the draws == the fit's own summary
    # 5. the FitOCT.R loop body in one call
    m = re.search(r"PIPELINE n_expgp (\d+) uy_len (\d+) mono_theta (\S+) (\S+) (\S+)", out)
    assert m and int(m.group(1)) == 3 and int(m.group(2)) == 3 * 481 and abs(float(m.group(5)) - 300) < 60
