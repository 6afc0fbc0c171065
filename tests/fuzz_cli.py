"""Randomised differential run of the drop-in CLI against the UNMODIFIED reference binary (test infrastructure; not part of the
product). For every case a small random data set is written as libFM text (two one-hot fields, or ragged rows with real values,
sometimes with comment / blank lines in the test file, sometimes converted to .x/.xt/.y with the tools), then
`oracle/_ref/libFM` (time() pinned by oracle/fixtime.c, as tests/golden/make_golden.py does) and `bin/libFM -seed <same>` are run
with the same arguments in two scratch directories, and the files they leave in the CWD (test_rmse_<tag>_<method>,
free_energy_<tag>_vb) and the `Train=` values on stdout are compared at north_star's 1e-4 (the files carry 6 digits).
vb / als / vb_online, -dim, -iter, -meta groups, -regular, -batch, -task c for als.

  SVBFM_EMU=1 python tests/fuzz_cli.py --seconds 300 --seed 1      # engine = tests/emu build (LD_PRELOAD), no GPU needed
  python tests/fuzz_cli.py --seconds 120                            # on a B200 (oracle/_ref travels with the snapshot)
"""
import argparse
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
BIN = os.path.join(ROOT, "scalable-variational-bayesian-factorization-machine_b200", "bin")
TOL = 1e-4


def write_two_field(path, n, U, I, r, binary=False):
    u, i = r.integers(0, U, n), r.integers(0, I, n)
    bu, bi = r.normal(0, 0.5, U), r.normal(0, 0.5, I)
    s = 3.5 + bu[u] + bi[i] + r.normal(0, 0.8, n)
    y = np.where(s > 3.5, 1, -1) if binary else np.clip(np.round(s), 1, 5).astype(int)
    with open(path, "w") as f:
        for a, b, c in zip(y, u, i):
            f.write(f"{a} {b}:1 {U + c}:1\n")
    return U + I


def write_ragged(path, n, D, r, comments):
    with open(path, "w") as f:
        if comments:
            f.write("# a comment line, then an empty line\n\n")
        for k in range(n):
            m = int(r.integers(0, 5))
            cols = np.sort(r.choice(D, size=min(m, D), replace=False))
            vals = np.round(r.uniform(0.2, 2.0, len(cols)), 3)
            y = round(float(r.normal(0, 1)), 3)
            lead = "  " if (comments and k % 7 == 0) else ""
            tail = "   # trailing" if (comments and k % 11 == 0) else ""
            f.write(lead + f"{y}" + "".join(f" {c}:{v}" for c, v in zip(cols, vals)) + tail + "\n")
    return D


def floats(path):
    return [float(x) for x in open(path).read().split()] if os.path.exists(path) else None


def train_values(stdout):
    return [float(l.split("Train=")[1].split("\t")[0]) for l in stdout.splitlines() if l.startswith("#Iter=") and "Train=" in l]


def close(a, b, tol=TOL):
    if a is None or b is None:
        return a is None and b is None
    if len(a) != len(b):
        return False
    for x, y in zip(a, b):
        if np.isnan(x) and np.isnan(y):
            continue
        if not abs(x - y) <= tol * max(abs(y), 1e-6) + 2e-6:       # 6 printed digits
            return False
    return True


def one_case(r, case_id, emu_env):
    kind = r.choice(["two", "two", "ragged"])
    method = r.choice(["vb", "vb", "als", "vb_online"])
    task_c = method == "als" and kind == "two" and r.random() < 0.3
    n, nt = int(r.choice([30, 400, 3000])), int(r.choice([10, 200]))
    K = int(r.choice([0, 1, 2, 4]))
    k0, k1 = int(r.random() < 0.8), int(r.random() < 0.8)
    iters = int(r.choice([1, 3, 6]))
    seed = int(r.integers(1, 100000))
    binary_files = r.random() < 0.3 and method != "vb_online"
    desc = f"case {case_id}: {kind} {method}{' -task c' if task_c else ''} n={n} nt={nt} dim={k0},{k1},{K} iter={iters} seed={seed} binary_files={binary_files}"
    with tempfile.TemporaryDirectory() as td:
        dirs = [os.path.join(td, "ref"), os.path.join(td, "my")]
        os.mkdir(dirs[0])
        if kind == "two":
            U, I = int(r.choice([4, 30, 200])), int(r.choice([3, 25, 150]))
            D = write_two_field(os.path.join(dirs[0], "train"), n, U, I, r, task_c)
            write_two_field(os.path.join(dirs[0], "test"), nt, U, I, r, task_c)
        else:
            D = int(r.choice([8, 40]))
            # the reference's vb_online maps raw file lines to batches (vbos.h:87-95): comment lines only in the test file
            write_ragged(os.path.join(dirs[0], "train"), n, D, r, comments=(method != "vb_online" and r.random() < 0.5))
            write_ragged(os.path.join(dirs[0], "test"), nt, D, r, comments=(r.random() < 0.5))
        args = ["-task", "c" if task_c else "r", "-train", "train", "-test", "test", "-dim", f"{k0},{k1},{K}", "-method", method, "-iter", str(iters)]
        if r.random() < 0.35:
            G = int(r.choice([2, 3]))
            with open(os.path.join(dirs[0], "meta"), "w") as f:       # one group id per attribute: ids < D plus the fork's phantom attribute
                f.write("\n".join(str(g) for g in (np.arange(D + 1) * G // (D + 1))) + "\n")
            args += ["-meta", "meta"]
            desc += f" groups={G}"
        if method == "als" and r.random() < 0.6:
            args += ["-regular", "0.1,0.5,1"]
            desc += " regular"
        if method == "vb_online":
            nb = int(r.choice([1, 2, 5]))
            if -(-n // nb) * (nb - 1) >= n:
                nb = 1
            args += ["-batch", str(nb)]
            desc += f" batch={nb}"
        if binary_files:
            for s in ("train", "test"):
                p = os.path.join(dirs[0], s)
                subprocess.run([os.path.join(BIN, "convert"), "--ifile", p, "--ofilex", p + ".x", "--ofiley", p + ".y"], check=True, capture_output=True)
                subprocess.run([os.path.join(BIN, "transpose"), "--ifile", p + ".x", "--ofile", p + ".xt"], check=True, capture_output=True)
        shutil.copytree(dirs[0], dirs[1])
        ref = subprocess.run([os.path.join(REF, "libFM")] + args, cwd=dirs[0], capture_output=True, text=True,
                             env=dict(os.environ, FAKE_TIME=str(seed), LD_PRELOAD=os.path.join(REF, "fixtime.so")))
        my = subprocess.run([os.path.join(BIN, "libFM")] + args + ["-seed", str(seed)], cwd=dirs[1], capture_output=True, text=True, env=emu_env)
        if "ERROR" in my.stderr or my.returncode != 0:
            if "ERROR" in ref.stderr or ref.returncode != 0:
                return desc + " (both refuse)", None
            return desc, "the CLI failed: " + my.stderr.strip()[-300:]
        if ref.returncode != 0:
            return desc + f" (the reference binary died with status {ref.returncode}; skipped)", None
        tag = f"{k0}{k1}{K}"
        m = "mcmc" if method == "als" else method
        names = [f"test_rmse_{tag}_{m}"] + ([f"free_energy_{tag}_vb"] if method in ("vb", "vb_online") else [])
        for name in names:
            a, b = floats(os.path.join(dirs[1], name)), floats(os.path.join(dirs[0], name))
            if not close(a, b):
                return desc, f"{name}: {a} != {b}"
        ta, tb = train_values(my.stdout), train_values(ref.stdout)
        if method != "vb_online" and not close(ta, tb):
            return desc, f"Train=: {ta} != {tb}"
    return desc, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cases", type=int, default=10**9)
    a = ap.parse_args()
    if not os.path.exists(os.path.join(REF, "libFM")):
        print("oracle/_ref/libFM is not built (make -C oracle ref needs /root/reference)")
        sys.exit(2)
    env = dict(os.environ)
    if os.environ.get("SVBFM_EMU"):
        sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
        import build_emu
        env["LD_PRELOAD"] = build_emu.build()
    r = np.random.default_rng(a.seed)
    t0, n = time.time(), 0
    while time.time() - t0 < a.seconds and n < a.cases:
        desc, err = one_case(r, n, env)
        print(desc, "OK" if err is None else "MISMATCH " + err, flush=True)
        if err is not None:
            sys.exit(1)
        n += 1
    print(f"{n} cases, no mismatch")


if __name__ == "__main__":
    main()
