"""CLI-level end-to-end time on the GPU box: file -> first iteration through bin/libFM, the way a
user of the reference runs it. Synthetic ratings of the ML-10M shape (bench generator) are written as libFM text, converted with
bin/convert, and `libFM -method vb -dim 1,1,K -iter 1` is timed by wall clock on (a) the text files (parallel parser + device
transpose), (b) the binary .x / .y files without .xt (device transpose), (c) with the .xt written by `bin/transpose --device 0`.
Prints one JSON line. Usage: python tools/cli_e2e.py [--rows N] [--k K]"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import svbfm_b200 as sv  # noqa: E402

BIN = os.path.join(ROOT, "scalable-variational-bayesian-factorization-machine_b200", "bin")


def write_text(path, u, i, y, U, I):
    """`y u:1 i:1` lines; the tokens come from per-user / per-item tables (object arrays: one string concatenation per line)."""
    ut = np.array([f" {k}:1" for k in range(U)], dtype=object)
    it = np.array([f" {U + k}:1" for k in range(I)], dtype=object)
    yv = y.astype(np.int64)
    yt = np.array([str(v) for v in range(int(yv.max()) + 1)], dtype=object)
    lines = yt[yv] + ut[u.astype(np.int64)] + it[i.astype(np.int64)]
    with open(path, "w") as f:
        f.write("\n".join(lines.tolist()))
        f.write("\n")


def timed(cmd, cwd):
    t0 = time.perf_counter()
    r = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    if r.returncode != 0 or "ERROR" in r.stderr:
        raise RuntimeError(" ".join(cmd) + "\n" + r.stdout[-1500:] + r.stderr[-1500:])
    return dt, r.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--k", type=int, default=50)
    a = ap.parse_args()
    synth = sv.submodule("synth")
    U, I, N, Nt, _ = synth.SHAPES["ml10m"]
    N, Nt = a.rows, max(1000, a.rows // 10)
    model = synth.planted_model(U, I, 7)
    u, i, y = synth.ratings(N, U, I, model, 1)
    ut, it, yt = synth.ratings(Nt, U, I, model, 2)
    out = {"workload": f"ml10m shape: {N} train / {Nt} test ratings, {U} users x {I} items, vb K={a.k}, 1 iteration", "host_threads": os.cpu_count()}
    with tempfile.TemporaryDirectory() as td:
        t0 = time.perf_counter()
        write_text(os.path.join(td, "tr.libfm"), u, i, y, U, I)
        write_text(os.path.join(td, "te.libfm"), ut, it, yt, U, I)
        out["write_text_s"] = round(time.perf_counter() - t0, 2)
        out["text_bytes"] = os.path.getsize(os.path.join(td, "tr.libfm")) + os.path.getsize(os.path.join(td, "te.libfm"))
        base = [os.path.join(BIN, "libFM"), "-task", "r", "-dim", f"1,1,{a.k}", "-method", "vb", "-iter", "1", "-seed", "42"]
        timed(base + ["-train", os.path.join(td, "te.libfm"), "-test", os.path.join(td, "te.libfm")], td)      # warm-up: CUDA context, page cache
        out["cli_text_s"], so = timed(base + ["-train", os.path.join(td, "tr.libfm"), "-test", os.path.join(td, "te.libfm")], td)
        out["test_rmse_text"] = float(open(os.path.join(td, f"test_rmse_11{a.k}_vb")).read().split()[-1])
        for n in ("tr", "te"):
            dt, _ = timed([os.path.join(BIN, "convert"), "--ifile", os.path.join(td, n + ".libfm"), "--ofilex", os.path.join(td, n + ".x"),
                           "--ofiley", os.path.join(td, n + ".y")], td)
            out[f"convert_{n}_s"] = round(dt, 2)
        out["cli_binary_rows_only_s"], _ = timed(base + ["-train", os.path.join(td, "tr"), "-test", os.path.join(td, "te")], td)
        out["test_rmse_binary"] = float(open(os.path.join(td, f"test_rmse_11{a.k}_vb")).read().split()[-1])
        for n in ("tr", "te"):
            dt, _ = timed([os.path.join(BIN, "transpose"), "--ifile", os.path.join(td, n + ".x"), "--ofile", os.path.join(td, n + ".xt"), "--device", "0"], td)
            out[f"transpose_device_{n}_s"] = round(dt, 2)
        out["cli_binary_with_xt_s"], _ = timed(base + ["-train", os.path.join(td, "tr"), "-test", os.path.join(td, "te")], td)
    for k in ("cli_text_s", "cli_binary_rows_only_s", "cli_binary_with_xt_s"):
        out[k] = round(out[k], 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
