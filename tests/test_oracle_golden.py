"""CPU: pins the oracle (oracle/svbfm_oracle.c) against the golden outputs of the UNMODIFIED reference binary
(tests/golden/golden.json, produced by tests/golden/make_golden.py with a fixed seed).

The reference prints 6 significant digits, so agreement is checked to 1e-5 relative: the oracle replays the
libc rand() stream and the expression order of the reference, and in practice matches every printed digit."""
import json
import os

import numpy as np
import pytest

import oracle_binding as ob

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLD = json.load(open(os.path.join(G, "golden.json")))["cases"]
# second set, CPU only: 100 iterations of the BASELINE config 1 switches (-dim '1,1,8'), mcmc without w0, als without -regular,
# vb_online with k0 = 0 and 7 batches, vb with groups and K = 1, vb on the ragged set
GOLD = GOLD + json.load(open(os.path.join(G, "golden_extra.json")))["cases"]
TOL = 1e-5


def load_case(c):
    tr = ob.parse_text(os.path.join(G, f"{c['data']}_train.libfm"))
    te = ob.parse_text(os.path.join(G, f"{c['data']}_test.libfm"))
    k0, k1, K = [int(x) for x in c["dim"].split(",")]
    kw = {}
    method = c["method"]
    if method == "als":
        method, kw = "mcmc", dict(do_sample=False, do_multilevel=False)
    extra = c.get("extra", [])
    if "-batch" in extra:
        kw["num_batch"] = int(extra[extra.index("-batch") + 1])
    if "-regular" in extra:
        kw["reg"] = [float(x) for x in extra[extra.index("-regular") + 1].split(",")]
    if c.get("meta"):
        kw["groups"] = np.loadtxt(os.path.join(G, "g2_meta.txt"), dtype=np.uint32)
    return tr, te, method, int(k0 != 0), int(k1 != 0), K, kw


@pytest.mark.parametrize("c", GOLD, ids=[c["name"] for c in GOLD])
def test_oracle_matches_reference_binary(c):
    tr, te, method, k0, k1, K, kw = load_case(c)
    orc = ob.Oracle(method, tr, te, K=K, k0=k0, k1=k1, seed=c["seed"], **kw)
    fe = []
    for it in range(c["iters"]):
        s = orc.iterate()
        assert abs(s.test_rmse - c["test_rmse"][it]) <= TOL * c["test_rmse"][it], (it, s.test_rmse, c["test_rmse"][it])
        if c["train_stat"]:
            assert abs(s.train_stat - c["train_stat"][it]) <= TOL * c["train_stat"][it]
        if s.has_free_energy:
            fe.append(-s.free_energy)
    if method == "vb":
        assert len(fe) == len(c["neg_free_energy"])
        for a, b in zip(fe, c["neg_free_energy"]):
            assert abs(a - b) <= TOL * abs(b), (a, b)
    if method == "vb_online":
        # the reference appends -F of batch 1 and batch B of every epoch; the oracle reports the last one per epoch
        for it, a in enumerate(fe):
            b = c["neg_free_energy"][2 * it + 1]
            assert abs(a - b) <= TOL * abs(b), (it, a, b)


def test_sa_fixture_when_reference_present():
    """The only real data the reference ships (data/sa.test_libfm): 90 000 / 10 000 split, vb '1,1,8'.
    Runs only where /root/reference and oracle/_ref exist (the authoring container)."""
    src = "/root/reference/data/sa.test_libfm"
    ref = os.path.join(os.path.dirname(G), "..", "oracle", "_ref", "libFM")
    if not (os.path.exists(src) and os.path.exists(ref)):
        pytest.skip("reference tree not present")
    import subprocess
    import tempfile
    lines = open(src).read().splitlines(True)
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "tr"), "w").writelines(lines[:90000])
        open(os.path.join(td, "te"), "w").writelines(lines[90000:])
        env = dict(os.environ, FAKE_TIME="42", LD_PRELOAD=os.path.join(os.path.dirname(ref), "fixtime.so"))
        subprocess.run([ref, "-task", "r", "-train", "tr", "-test", "te", "-dim", "1,1,8", "-method", "vb", "-iter", "5"], cwd=td, env=env,
                       check=True, capture_output=True)
        want = [float(x) for x in open(os.path.join(td, "test_rmse_118_vb")).read().split()]
        want_fe = [float(x) for x in open(os.path.join(td, "free_energy_118_vb")).read().split()]
        tr, te = ob.parse_text(os.path.join(td, "tr")), ob.parse_text(os.path.join(td, "te"))
    orc = ob.Oracle("vb", tr, te, K=8, seed=42)
    for it in range(5):
        s = orc.iterate()
        assert abs(s.test_rmse - want[it]) <= TOL * want[it]
        assert abs(-s.free_energy - want_fe[it]) <= TOL * abs(want_fe[it])


GOLD_CLASS = json.load(open(os.path.join(G, "golden_class.json")))["cases"]


def load_class_case(c):
    tr, te, method, k0, k1, K, kw = load_case(c)
    for s in (tr, te):          # libfm.cpp:339-340: every target <= 0 becomes -1, the others +1
        s.y[:] = np.where(s.y <= 0.0, -1.0, 1.0).astype(np.float32)
    return tr, te, method, k0, k1, K, kw


@pytest.mark.parametrize("c", GOLD_CLASS, ids=[c["name"] for c in GOLD_CLASS])
def test_oracle_classification_matches_reference_binary(c):
    """-task c (mcmc / als): train and test accuracy per iteration as the unmodified reference binary prints them, including
    the truncated-normal target draws on the libc stream (util/random.h:72-114) and the reference's own erf."""
    tr, te, method, k0, k1, K, kw = load_class_case(c)
    orc = ob.Oracle(method, tr, te, K=K, k0=k0, k1=k1, seed=c["seed"], task=1, **kw)
    for it in range(c["iters"]):
        s = orc.iterate()
        assert abs(s.train_stat - c["train_stat"][it]) <= TOL * c["train_stat"][it], (it, s.train_stat, c["train_stat"][it])
        assert abs(s.test_rmse - c["test_acc"][it]) <= TOL * c["test_acc"][it], (it, s.test_rmse, c["test_acc"][it])
