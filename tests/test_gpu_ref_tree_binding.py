"""-m gpu: the literal drop-in proof. oracle/_ref/libFM_cuda is the REFERENCE's own libfm.cpp compiled with the binding headers
host/reference_tree/fm_learn_{vb,mcmc}_cuda.h against libsvbfm.so (recipe: oracle/make_ref_cuda.py). One binary, same seed
(the time() shim, because the reference seeds with time(NULL)): `-method vb` is the unmodified reference learner,
`-method vb_cuda` the reference's class fm_learn_vb (its init(), its fields) with the sweep on the B200. What both leave in
the CWD and print is compared at north_star's 1e-4 (the files carry 6 significant digits)."""
import os
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
REF = os.path.join(ROOT, "oracle", "_ref")
EXE = os.path.join(REF, "libFM_cuda")
TOL = 1e-4


def floats(path):
    return [float(x) for x in open(path).read().split()]


def run(tmp, method, data, dim, iters, extra=(), seed=42):
    d = tmp / method
    d.mkdir()
    for s in ("train", "test"):
        shutil.copy(os.path.join(G, f"{data}_{s}.libfm"), d / s)
    if os.path.exists(os.path.join(G, "g2_meta.txt")):
        shutil.copy(os.path.join(G, "g2_meta.txt"), d / "meta")
    pre = [os.path.join(REF, "fixtime.so")]
    if os.environ.get("SVBFM_EMU_PRELOAD"):      # the engine built for tests/emu (CPU suite: checks the glue without a GPU)
        pre.append(os.environ["SVBFM_EMU_PRELOAD"])
    env = dict(os.environ, FAKE_TIME=str(seed), LD_PRELOAD=":".join(pre))
    p = subprocess.run([EXE, "-task", "r", "-train", "train", "-test", "test", "-dim", dim, "-method", method, "-iter", str(iters), "-out", "pred.txt"] + list(extra),
                       cwd=d, env=env, capture_output=True, text=True, timeout=600)
    assert "ERROR" not in p.stderr, p.stderr
    return d, p.stdout


def close(a, b, tol=TOL):
    assert len(a) == len(b) and len(a) > 0, (len(a), len(b))
    for k, (x, y) in enumerate(zip(a, b)):
        assert abs(x - y) <= tol * abs(y), (k, x, y)


CASES = [("g1", "1,1,4", 10, ()), ("g1", "1,1,8", 30, ()), ("g2", "1,1,3", 8, ("-meta", "meta")), ("g3", "0,1,2", 6, ())]


@pytest.mark.parametrize("data,dim,iters,extra", CASES, ids=[f"{c[0]}-{c[1]}" for c in CASES])
def test_vb_cuda_equals_vb_in_the_reference_binary(tmp_path, data, dim, iters, extra):
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/libFM_cuda not built (needs /root/reference at build time)")
    k = dim.split(",")
    tag = f"{int(k[0] != '0')}{int(k[1] != '0')}{k[2]}"
    d_ref, out_ref = run(tmp_path, "vb", data, dim, iters, extra)
    d_gpu, out_gpu = run(tmp_path, "vb_cuda", data, dim, iters, extra)
    assert "in learn of fm_learn_vb_cuda" in out_gpu and "in learn of fm_learn_vb_cuda" not in out_ref
    close(floats(d_gpu / f"test_rmse_{tag}_vb"), floats(d_ref / f"test_rmse_{tag}_vb"))
    close(floats(d_gpu / f"free_energy_{tag}_vb"), floats(d_ref / f"free_energy_{tag}_vb"))
    tr = lambda o: [float(l.split("Train=")[1].split("\t")[0]) for l in o.splitlines() if l.startswith("#Iter=")]
    close(tr(out_gpu), tr(out_ref))
    assert len(floats(d_gpu / "pred.txt")) == sum(1 for _ in open(d_gpu / "test"))


def test_als_cuda_equals_als_in_the_reference_binary(tmp_path):
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/libFM_cuda not built (needs /root/reference at build time)")
    d_ref, out_ref = run(tmp_path, "als", "g1", "1,1,4", 8, ("-regular", "0,1,2"))
    d_gpu, out_gpu = run(tmp_path, "als_cuda", "g1", "1,1,4", 8, ("-regular", "0,1,2"))
    close(floats(d_gpu / "test_rmse_114_mcmc"), floats(d_ref / "test_rmse_114_mcmc"))
    close(floats(d_gpu / "pred.txt"), floats(d_ref / "pred.txt"), 2e-4)
