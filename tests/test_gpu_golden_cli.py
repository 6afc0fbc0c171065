"""-m gpu: the drop-in path end to end. The libFM-compatible CLI (host/libfm.cpp -> C-ABI -> CUDA) is run on the
committed golden inputs with the same seed the reference binary was run with, and the files it writes into the CWD
(test_rmse_<k0k1K>_<method>, free_energy_<k0k1K>_vb) are compared with the reference's own files.
Tolerance: north_star's 1e-4 relative per iteration for VB / ALS (values carry 6 significant digits); sampled MCMC
uses another RNG stream and is only checked for sanity here (distribution parity: test_gpu_parity.py)."""
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

import svbfm_b200 as sv

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
GOLD = json.load(open(os.path.join(G, "golden.json")))["cases"]
# six more runs of the reference binary, among them 100 iterations of `-method vb -dim '1,1,8'` (the switches of BASELINE config 1)
GOLD = GOLD + json.load(open(os.path.join(G, "golden_extra.json")))["cases"]
EXE = os.path.join(sv.PKG_DIR, "bin", "libFM")
TOL = 1e-4


def floats(path):
    return [float(x) for x in open(path).read().split()]


@pytest.mark.parametrize("c", GOLD, ids=lambda c: c["name"])
def test_cli_reproduces_reference_files(built, tmp_path, c):
    for s in ("train", "test"):
        shutil.copy(os.path.join(G, f"{c['data']}_{s}.libfm"), tmp_path / s)
    args = [EXE, "-task", "r", "-train", "train", "-test", "test", "-dim", c["dim"], "-method", c["method"], "-iter", str(c["iters"]),
            "-seed", str(c["seed"]), "-rlog", "log.tsv", "-out", "pred.txt"]
    if c.get("meta"):
        shutil.copy(os.path.join(G, "g2_meta.txt"), tmp_path / "meta")
        args += ["-meta", "meta"]
    args += c.get("extra", [])
    p = subprocess.run(args, cwd=tmp_path, capture_output=True, text=True)
    assert "ERROR" not in p.stderr, p.stderr
    k = c["dim"].split(",")
    tag = f"{int(k[0] != '0')}{int(k[1] != '0')}{k[2]}"
    m = "mcmc" if c["method"] == "als" else c["method"]
    rmse = floats(tmp_path / f"test_rmse_{tag}_{m}")
    assert len(rmse) == c["iters"]
    exact = c["method"] in ("vb", "als", "vb_online")
    if exact:
        for it, (a, b) in enumerate(zip(rmse, c["test_rmse"])):
            assert abs(a - b) <= TOL * b, (it, a, b)
        train = [float(l.split("Train=")[1].split("\t")[0]) for l in p.stdout.splitlines() if l.startswith("#Iter=") and "Train=" in l]
        for a, b in zip(train, c["train_stat"]):
            assert abs(a - b) <= TOL * b
    else:
        assert abs(rmse[-1] - c["test_rmse"][-1]) < 0.05 * c["test_rmse"][-1]
    if c["method"] in ("vb", "vb_online"):     # vb_online appends batch 1 and batch B of every epoch to the _vb file (vbo.h:637)
        fe = floats(tmp_path / f"free_energy_{tag}_vb")
        assert len(fe) == len(c["neg_free_energy"])
        for a, b in zip(fe, c["neg_free_energy"]):
            assert abs(a - b) <= TOL * abs(b), (a, b)
    pred = floats(tmp_path / "pred.txt")
    assert len(pred) == sum(1 for _ in open(tmp_path / "test")) and np.all(np.isfinite(pred))
    log = [l.split("\t") for l in open(tmp_path / "log.tsv").read().splitlines()]
    assert "time_learn" in log[0] and len(log) == c["iters"] + 1


def test_cli_binary_input_equals_text_input(built, tmp_path):
    """`-train name` with name.x / name.xt / name.y next to it (Data.h:112-117) gives the same numbers as the text file."""
    c = GOLD[0]
    BIN = os.path.join(sv.PKG_DIR, "bin")
    for s in ("train", "test"):
        src = os.path.join(G, f"{c['data']}_{s}.libfm")
        subprocess.run([os.path.join(BIN, "convert"), "--ifile", src, "--ofilex", str(tmp_path / f"{s}.x"), "--ofiley", str(tmp_path / f"{s}.y")],
                       check=True, capture_output=True)
        subprocess.run([os.path.join(BIN, "transpose"), "--ifile", str(tmp_path / f"{s}.x"), "--ofile", str(tmp_path / f"{s}.xt")], check=True, capture_output=True)
    p = subprocess.run([EXE, "-task", "r", "-train", "train", "-test", "test", "-dim", c["dim"], "-method", "vb", "-iter", "3", "-seed", str(c["seed"])],
                       cwd=tmp_path, capture_output=True, text=True)
    assert "ERROR" not in p.stderr, p.stderr
    rmse = floats(tmp_path / "test_rmse_114_vb")
    for a, b in zip(rmse, c["test_rmse"][:3]):
        assert abs(a - b) <= TOL * b


def test_cli_baseline_config_1_on_the_reference_data(built, tmp_path):
    """BASELINE config 1: the reference's own data (data/sa.test_libfm, committed gzipped; 90 000 / 10 000 split), `-method vb -dim
    '1,1,8' -iter 100`: every iteration's test RMSE, train RMSE and free energy against the files the unmodified reference binary wrote
    (tests/golden/make_golden_sa.py), 1e-4 relative."""
    import gzip
    c = json.load(open(os.path.join(G, "golden_sa.json")))
    lines = gzip.open(os.path.join(G, "sa.test_libfm.gz"), "rt").read().splitlines(True)
    assert len(lines) == 100000
    open(tmp_path / "tr", "w").writelines(lines[:90000])
    open(tmp_path / "te", "w").writelines(lines[90000:])
    p = subprocess.run([EXE, "-task", "r", "-train", "tr", "-test", "te", "-dim", "1,1,8", "-method", "vb", "-iter", str(c["iters"]), "-seed", str(c["seed"])],
                       cwd=tmp_path, capture_output=True, text=True)
    assert "ERROR" not in p.stderr, p.stderr
    rmse, fe = floats(tmp_path / "test_rmse_118_vb"), floats(tmp_path / "free_energy_118_vb")
    train = [float(l.split("Train=")[1].split("\t")[0]) for l in p.stdout.splitlines() if l.startswith("#Iter=")]
    assert len(rmse) == len(fe) == len(train) == c["iters"]
    for it in range(c["iters"]):
        assert abs(rmse[it] - c["test_rmse"][it]) <= TOL * c["test_rmse"][it], (it, rmse[it], c["test_rmse"][it])
        assert abs(train[it] - c["train_stat"][it]) <= TOL * c["train_stat"][it], (it, train[it], c["train_stat"][it])
        assert abs(fe[it] - c["neg_free_energy"][it]) <= TOL * abs(c["neg_free_energy"][it]), (it, fe[it], c["neg_free_energy"][it])


def test_cli_save_and_load_model_resume_bit_for_bit(built, tmp_path):
    """-save_model / -load_model: 5 iterations, save, load, 5 more iterations write the same lines into
    test_rmse_* and free_energy_* as 10 iterations in one go: parameters, hyper-parameters, residuals and sum T travel as raw doubles."""
    c = GOLD[0]
    for s in ("train", "test"):
        shutil.copy(os.path.join(G, f"{c['data']}_{s}.libfm"), tmp_path / s)
    base = [EXE, "-task", "r", "-train", "train", "-test", "test", "-dim", c["dim"], "-method", "vb", "-seed", str(c["seed"])]
    k = c["dim"].split(",")
    tag = f"{int(k[0] != '0')}{int(k[1] != '0')}{k[2]}"

    def run(extra):
        p = subprocess.run(base + extra, cwd=tmp_path, capture_output=True, text=True)
        assert "ERROR" not in p.stderr, p.stderr
        return open(tmp_path / f"test_rmse_{tag}_vb").read().split(), open(tmp_path / f"free_energy_{tag}_vb").read().split()

    full_r, full_f = run(["-iter", "10"])
    a_r, a_f = run(["-iter", "5", "-save_model", "model.bin"])
    b_r, b_f = run(["-iter", "5", "-load_model", "model.bin", "-out", "pred.txt"])
    assert len(full_r) == 10 and a_r + b_r == full_r and a_f + b_f == full_f
    # a model of other dimensions is refused the way the reference reports errors
    other = [EXE, "-task", "r", "-train", "train", "-test", "test", "-dim", "1,1,2", "-method", "vb", "-iter", "1", "-load_model", "model.bin"]
    p = subprocess.run(other, cwd=tmp_path, capture_output=True, text=True)
    assert "ERROR" in p.stderr and "dimensions" in p.stderr


@pytest.mark.parametrize("name", ["g1_train", "g2_train"])
def test_transpose_tool_on_the_device_writes_the_reference_bytes(built, tmp_path, name):
    """`transpose --device 0` (svbfm_transpose_csr: entry -> case id, one stable radix sort by feature id on the
    GPU) writes the .xt the unmodified reference tool wrote (tests/golden/*.xt), byte for byte."""
    BIN = os.path.join(sv.PKG_DIR, "bin")
    env = dict(os.environ, SVBFM_LIB=sv.LIB_PATH)
    p = subprocess.run([os.path.join(BIN, "transpose"), "--ifile", os.path.join(G, name + ".x"), "--ofile", str(tmp_path / "out.xt"), "--device", "0"],
                       capture_output=True, text=True, env=env)
    assert p.returncode == 0 and not p.stderr.strip(), p.stderr
    assert open(tmp_path / "out.xt", "rb").read() == open(os.path.join(G, name + ".xt"), "rb").read()


def test_cli_rows_only_binary_input_is_transposed_on_the_device(built, tmp_path):
    """`-train name` with only name.x + name.y next to it (no name.xt): the CLI hands the rows to svbfm_set_csr and the device builds the
    transposed matrix; text input takes the same route. Same numbers as the reference's run on the text file; with SVBFM_HOST_TRANSPOSE=1
    (host transpose + svbfm_set_csc, the round-1 path) the files are identical."""
    c = GOLD[0]
    BIN = os.path.join(sv.PKG_DIR, "bin")
    for s in ("train", "test"):
        src = os.path.join(G, f"{c['data']}_{s}.libfm")
        shutil.copy(src, tmp_path / (s + ".txt"))
        subprocess.run([os.path.join(BIN, "convert"), "--ifile", src, "--ofilex", str(tmp_path / f"{s}.x"), "--ofiley", str(tmp_path / f"{s}.y")],
                       check=True, capture_output=True)
    base = ["-task", "r", "-dim", c["dim"], "-method", "vb", "-iter", "3", "-seed", str(c["seed"])]
    outs = []
    for train, test, env in (("train", "test", {}), ("train.txt", "test.txt", {}), ("train.txt", "test.txt", {"SVBFM_HOST_TRANSPOSE": "1"})):
        p = subprocess.run([EXE, "-train", train, "-test", test] + base, cwd=tmp_path, capture_output=True, text=True, env=dict(os.environ, **env))
        assert "ERROR" not in p.stderr, p.stderr
        outs.append((open(tmp_path / "test_rmse_114_vb").read(), open(tmp_path / "free_energy_114_vb").read()))
    assert outs[0] == outs[1] == outs[2]
    for a, b in zip([float(x) for x in outs[0][0].split()], c["test_rmse"][:3]):
        assert abs(a - b) <= TOL * b
