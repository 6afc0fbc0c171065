#!/usr/bin/env python
"""Generates tests/golden/sa.test_libfm.gz + golden_sa.json: BASELINE config 1 on the only data set the reference ships
(data/sa.test_libfm, 100 000 ratings; sa.train_libfm itself is missing from the reference): the first 90 000 lines are the train split,
the last 10 000 the test split, and the UNMODIFIED reference binary (oracle/_ref/libFM, seed pinned by oracle/fixtime.c) runs
`-method vb -dim '1,1,8' -iter 100` on them. Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden_sa.py
"""
import gzip
import json
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")
SRC = "/root/reference/data/sa.test_libfm"
SEED = 42


def main():
    raw = open(SRC, "rb").read()
    with open(os.path.join(HERE, "sa.test_libfm.gz"), "wb") as f:
        f.write(gzip.compress(raw, 9, mtime=0))
    lines = raw.decode().splitlines(True)
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "tr"), "w").writelines(lines[:90000])
        open(os.path.join(td, "te"), "w").writelines(lines[90000:])
        env = dict(os.environ, FAKE_TIME=str(SEED), LD_PRELOAD=os.path.join(REF, "fixtime.so"))
        p = subprocess.run([os.path.join(REF, "libFM"), "-task", "r", "-train", "tr", "-test", "te", "-dim", "1,1,8", "-method", "vb", "-iter", "100"],
                           cwd=td, env=env, check=True, capture_output=True, text=True)
        floats = lambda name: [float(x) for x in open(os.path.join(td, name)).read().split()]
        out = {"generator": "tests/golden/make_golden_sa.py", "data": "sa.test_libfm.gz (reference data/sa.test_libfm): lines 1-90000 train, 90001-100000 test",
               "command": "-task r -dim 1,1,8 -method vb -iter 100", "seed": SEED, "iters": 100,
               "test_rmse": floats("test_rmse_118_vb"), "neg_free_energy": floats("free_energy_118_vb"),
               "train_stat": [float(l.split("Train=")[1].split("\t")[0]) for l in p.stdout.splitlines() if l.startswith("#Iter=")]}
    json.dump(out, open(os.path.join(HERE, "golden_sa.json"), "w"), indent=0)
    print("wrote", len(out["test_rmse"]), "iterations; last test rmse", out["test_rmse"][-1])


if __name__ == "__main__":
    main()
