"""Variants of the hot path behind environment knobs against the oracle: the layouts that are NOT the default any more (records in
column order, SVBFM_REC_RANK=0), the masked vb_online passes, graph replay, binary classification. Kept in a file that sorts
last: the default path's parity tests come first."""
import numpy as np
import pytest

import oracle_binding as ob
from helpers import make_learner, rel, to_csc, two_field

pytestmark = pytest.mark.gpu
VB_TOL = 1e-7


@pytest.mark.parametrize("tile_entries", [0, 64])
@pytest.mark.parametrize("rank", ["1", "0"])
def test_rec_rank_layout(built, monkeypatch, tile_entries, rank):
    """Rank layout (default since round 2; SVBFM_REC_RANK=0: records in column order): second-field records by popularity rank,
    cases of a first-field column ordered by that rank. Same algorithm, another case order: statistics, parameters and residuals
    (caller order) equal the oracle's either way."""
    monkeypatch.setenv("SVBFM_REC_RANK", rank)
    bits = 3 if rank == "1" else 1
    for values in (False, True):
        tr, te = two_field(20000, 2000, 300, 200, seed=11, values=values)
        orc = ob.Oracle("vb", tr, te, K=3, seed=42)
        L = make_learner("vb", tr, te, 3, num_iter=4, tile_entries=tile_entries)
        for it, s in enumerate(L.learn(to_csc(tr), to_csc(te))):
            o = orc.iterate()
            assert rel(s.test_rmse, o.test_rmse) < VB_TOL and rel(s.free_energy, o.free_energy) < VB_TOL and rel(s.alpha, o.alpha) < VB_TOL, it
        assert L.engine.info()["fused_schedule"] & 3 == bits
        assert L.engine.copies_max_diff() == 0.0
        e_o, t_o = orc.get_train_cache()
        assert np.max(np.abs(L.engine.get_residuals() - e_o)) < (1e-7 if values else 1e-9)   # x != 1: float products round differently
        so, sg = orc.get_state(), L.engine.get_state()
        for k in ("w_mean", "w_var", "v_mean", "v_var"):
            assert np.max(np.abs(so[k] - sg[k])) < (1e-7 if values else 1e-9), k
        L.engine.close()
    # als and vb_online on the same layout
    tr, te = two_field(12000, 1200, 200, 150, seed=12)
    orc = ob.Oracle("mcmc", tr, te, K=2, seed=42, do_sample=False, do_multilevel=False)
    L = make_learner("mcmc", tr, te, 2, num_iter=3, do_sample=False, do_multilevel=False, tile_entries=tile_entries)
    L.fm.reg0, L.fm.regw, L.fm.regv = 0.0, 0.0, 0.0
    for s in L.learn(to_csc(tr), to_csc(te)):
        o = orc.iterate()
        assert rel(s.test_rmse, o.test_rmse) < VB_TOL and rel(s.train_stat, o.train_stat) < VB_TOL
    assert L.engine.info()["fused_schedule"] & 3 == bits
    L.engine.close()
    orc = ob.Oracle("vb_online", tr, te, K=2, seed=42, num_batch=5)
    want = [orc.iterate() for _ in range(2)]
    L = make_learner("vb_online", tr, te, 2, num_iter=2, num_batch=5, tile_entries=tile_entries)
    for it, s in enumerate(L.learn(to_csc(tr), to_csc(te))):
        assert rel(s.test_rmse, want[it].test_rmse) < VB_TOL and rel(s.free_energy, want[it].free_energy) < VB_TOL
    assert L.engine.info()["fused_schedule"] & 3 == bits
    L.engine.close()


def test_vb_online_batch_lists_equal_masked_passes(built, monkeypatch):
    """vb_online on the stream schedule: prediction / reductions / w0 shift over the batch's own case list (default) against
    the masked passes over the whole arrays (SVBFM_VBO_FULL_PASSES=1): same numbers up to the order of the sums."""
    tr, te = two_field(16000, 1600, 260, 190, seed=21)
    out = []
    for full in (False, True):
        if full:
            monkeypatch.setenv("SVBFM_VBO_FULL_PASSES", "1")
        else:
            monkeypatch.delenv("SVBFM_VBO_FULL_PASSES", raising=False)
        L = make_learner("vb_online", tr, te, 3, num_iter=3, num_batch=7, tile_entries=64)
        out.append([(s.test_rmse, s.free_energy, s.alpha) for s in L.learn(to_csc(tr), to_csc(te))])
        assert L.engine.info()["fused_schedule"] & 1
        L.engine.close()
    for a, b in zip(*out):
        assert all(rel(x, y) < 1e-10 for x, y in zip(a, b)), (a, b)


@pytest.mark.parametrize("values,tile_entries", [(False, 0), (False, 64), (True, 64), (False, 32)])
def test_vb_online_packed_batches_equal_index_lists(built, monkeypatch, values, tile_entries):
    """vb_online on the stream schedule: the batch in flight packed into contiguous streams against passes that read through the
    batch's index lists (SVBFM_VBO_PACK=0). Packed and swept by k_stream (SVBFM_VBO_ROWS=0; all-ones data then goes through the
    bulk-copy ring like a whole-run pass): same operands in the same order, so statistics, parameters and residuals are
    bit-identical. Packed and swept by k_stream_rows (the default): the columns that begin and end inside a row of 32 entries
    are summed by a segmented scan instead of the butterfly, everything else is the same arithmetic: equal to rounding.
    SVBFM_VBO_COMPACT=0 keeps global column ids in k_stream_rows / k_finalize_vbo."""
    tr, te = two_field(16000, 1600, 260, 190, seed=22, values=values)
    out, res, par = [], [], []
    for packed, rows, compact in (("0", "0", "1"), ("1", "0", "1"), ("1", "1", "1"), ("1", "1", "0")):
        monkeypatch.setenv("SVBFM_VBO_PACK", packed)
        monkeypatch.setenv("SVBFM_VBO_ROWS", rows)
        monkeypatch.setenv("SVBFM_VBO_COMPACT", compact)
        L = make_learner("vb_online", tr, te, 3, num_iter=3, num_batch=7, tile_entries=tile_entries)
        out.append([(s.test_rmse, s.free_energy, s.alpha) for s in L.learn(to_csc(tr), to_csc(te))])
        assert L.engine.info()["fused_schedule"] & 1
        res.append(L.engine.get_residuals())
        par.append(np.concatenate([np.ravel(x) for x in L.engine.get_state().values()]))
        assert L.engine.copies_max_diff() == 0.0
        L.engine.close()
    assert out[0] == out[1], out
    assert np.array_equal(res[0], res[1]) and np.array_equal(par[0], par[1])
    for a, b in zip(out[0], out[2]):
        assert all(rel(x, y) < 1e-10 for x, y in zip(a, b)), (a, b)
    assert np.max(np.abs(res[0] - res[2])) < 1e-9 and np.max(np.abs(par[0] - par[2])) < 1e-9
    assert not np.array_equal(par[0], par[2]) or tile_entries == 0      # (the rows kernel did run: its sums round differently)
    # the batch's non-empty columns as a dense id space (default on one GPU) against global column ids: the addresses differ, and
    # d(sum T) of a batch is summed over the dense array instead of over all columns (the same values between other zeros)
    for a, b in zip(out[2], out[3]):
        assert all(rel(x, y) < 1e-12 for x, y in zip(a, b)), (a, b)
    assert np.max(np.abs(res[2] - res[3])) < 1e-11 and np.max(np.abs(par[2] - par[3])) < 1e-11


# ---- binary classification (-task c) for mcmc / als: SURVEY section 8(f) rank 4. Built and emulator-checked without a GPU.
def _binary_two_field(n, nt, U, I, seed, values=False):
    tr, te = two_field(n, nt, U, I, seed=seed, values=values)
    for s in (tr, te):
        s.y[:] = np.where(s.y >= 4, 1.0, -1.0).astype(np.float32)     # libfm.cpp:339-340 maps to -1 / +1
    return tr, te


@pytest.mark.parametrize("values,tile_entries", [(False, 0), (True, 64)])
def test_classification_als_equals_oracle(built, values, tile_entries):
    """do_sample = 0: the latent targets are the expected values of the truncated normals (mcmcs.h:199-216): deterministic,
    so accuracies, residuals (yhat - latent target) and the averaged test probabilities equal the oracle's."""
    tr, te = _binary_two_field(12000, 1500, 200, 150, seed=51, values=values)
    orc = ob.Oracle("mcmc", tr, te, K=3, seed=42, do_sample=False, do_multilevel=False, reg=(0.0, 0.5, 1.0), task=1)
    L = make_learner("mcmc", tr, te, 3, num_iter=6, do_sample=False, do_multilevel=False, tile_entries=tile_entries, task=1)
    L.fm.reg0, L.fm.regw, L.fm.regv = 0.0, 0.5, 1.0
    for it, s in enumerate(L.learn(to_csc(tr), to_csc(te))):
        o = orc.iterate()
        assert abs(s.train_stat - o.train_stat) < 2.0 / tr.n_rows, (it, s.train_stat, o.train_stat)     # at most a borderline case or two
        assert abs(s.test_rmse - o.test_rmse) < 2.0 / te.n_rows and abs(s.rmse_this - o.rmse_this) < 2.0 / te.n_rows
    e_o, _ = orc.get_train_cache(want_t=False)
    assert np.max(np.abs(L.engine.get_residuals() - e_o)) < 1e-6
    assert np.max(np.abs(L.engine.predict() - orc.get_test_pred())) < 1e-6
    if not values:
        assert L.engine.info()["fused_schedule"] & 1 and L.engine.copies_max_diff() == 0.0
    L.engine.close()


def test_classification_golden_als(built):
    """The als run of the reference binary on tests/golden/g4_* (-task c): same accuracies as printed by the reference."""
    import json
    import os
    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    c = [c for c in json.load(open(os.path.join(G, "golden_class.json")))["cases"] if c["name"] == "g4_als_c_112"][0]
    tr, te = ob.parse_text(os.path.join(G, "g4_train.libfm")), ob.parse_text(os.path.join(G, "g4_test.libfm"))
    lo, hi = float(tr.y.min()), float(tr.y.max())
    for s in (tr, te):
        s.y[:] = np.where(s.y <= 0.0, -1.0, 1.0).astype(np.float32)
    reg = [float(x) for x in c["extra"][1].split(",")]
    L = make_learner("mcmc", tr, te, 2, seed=c["seed"], num_iter=c["iters"], do_sample=False, do_multilevel=False, task=1)
    L.min_target, L.max_target = lo, hi
    L.fm.reg0, L.fm.regw, L.fm.regv = reg
    for it, s in enumerate(L.learn(to_csc(tr), to_csc(te))):
        assert abs(s.train_stat - c["train_stat"][it]) < 1e-5 + 1.0 / tr.n_rows, (it, s.train_stat, c["train_stat"][it])
        assert abs(s.test_rmse - c["test_acc"][it]) < 1e-5 + 1.0 / te.n_rows, (it, s.test_rmse, c["test_acc"][it])
    L.engine.close()


def test_classification_sampling_distribution(built):
    """do_sample = 1: Philox draws instead of the libc stream, so only the distribution can match: train / test accuracy of
    the running mean after 12 sweeps within 2 points of the mean over three oracle seeds."""
    tr, te = _binary_two_field(20000, 4000, 150, 100, seed=52)
    want_tr, want_te = [], []
    for seed in (1, 2, 3):
        orc = ob.Oracle("mcmc", tr, te, K=3, seed=seed, task=1)
        for _ in range(12):
            o = orc.iterate()
        want_tr.append(o.train_stat); want_te.append(o.test_rmse)
    L = make_learner("mcmc", tr, te, 3, num_iter=12, task=1)
    s = L.learn(to_csc(tr), to_csc(te))[-1]
    assert abs(s.train_stat - np.mean(want_tr)) < 0.02, (s.train_stat, want_tr)
    assert abs(s.test_rmse - np.mean(want_te)) < 0.02, (s.test_rmse, want_te)
    assert 0.5 < s.test_rmse <= 1.0
    p = L.engine.predict()
    assert p.min() >= 0.0 and p.max() <= 1.0
    L.engine.close()


def test_cli_classification_reproduces_reference_stdout(built, tmp_path):
    """libFM-compatible CLI with -task c -method als on tests/golden/g4_*: the '#Iter= .. Train= .. Test= .. MAP@5= 0' lines of
    the unmodified reference binary (nothing is written to test_rmse_* for classification), and -out probabilities in [0, 1]."""
    import json
    import os
    import shutil
    import subprocess
    import svbfm_b200 as sv
    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    c = [c for c in json.load(open(os.path.join(G, "golden_class.json")))["cases"] if c["name"] == "g4_als_c_112"][0]
    for s in ("train", "test"):
        shutil.copy(os.path.join(G, f"g4_{s}.libfm"), tmp_path / s)
    args = [os.path.join(sv.PKG_DIR, "bin", "libFM"), "-task", "c", "-train", "train", "-test", "test", "-dim", c["dim"], "-method", "als",
            "-iter", str(c["iters"]), "-seed", str(c["seed"]), "-out", "pred.txt"] + c["extra"]
    p = subprocess.run(args, cwd=tmp_path, capture_output=True, text=True)
    assert "ERROR" not in p.stderr, p.stderr
    lines = [l for l in p.stdout.splitlines() if l.startswith("#Iter=")]
    assert len(lines) == c["iters"] and all(l.endswith("MAP@5= 0") for l in lines)
    for it, l in enumerate(lines):
        tr_acc, te_acc = float(l.split("Train=")[1].split("\t")[0]), float(l.split("Test=")[1].split("\t")[0])
        assert abs(tr_acc - c["train_stat"][it]) < 1e-4 + 1.0 / 2500 and abs(te_acc - c["test_acc"][it]) < 1e-4 + 1.0 / 400, (it, l)
    assert open(tmp_path / "test_rmse_112_mcmc").read() == ""
    pred = [float(x) for x in open(tmp_path / "pred.txt").read().split()]
    assert len(pred) == 400 and min(pred) >= 0.0 and max(pred) <= 1.0


@pytest.mark.parametrize("method", ["vb", "mcmc"])
def test_graph_replay_equals_plain_launches(built, monkeypatch, method):
    """SVBFM_GRAPH=1: iterations 1.. of svbfm_run are replays of a CUDA graph captured from one iteration. Same launches with
    the same arguments in the same order: every statistic is bit-identical to the plain run, and so are the parameters."""
    tr, te = two_field(15000, 1500, 250, 180, seed=61)
    kw = dict(do_sample=True, do_multilevel=True) if method == "mcmc" else {}
    out = []
    for graph in ("0", "1"):
        monkeypatch.setenv("SVBFM_GRAPH", graph)
        L = make_learner(method, tr, te, 3, num_iter=5, **kw)
        hist = L.learn(to_csc(tr), to_csc(te))
        out.append([(s.test_rmse, s.train_stat, s.free_energy, s.alpha, s.rmse_this, s.nan_inf_count) for s in hist])
        st = L.engine.get_state()
        out.append(np.concatenate([st["w_mean"], st["v_mean"].ravel(), st["v_var"].ravel()]))
        assert bool(L.engine.info()["fused_schedule"] & 8) == (graph == "1")
        assert L.engine.info()["kernel_launches"] > 0
        out.append(L.engine.info()["kernel_launches"])
        L.engine.close()
    assert out[0] == out[3]
    assert np.array_equal(out[1], out[4])
    assert out[2] == out[5]          # the launch counter counts the replayed kernels too
