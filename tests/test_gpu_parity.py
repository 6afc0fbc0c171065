"""-m gpu: the CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs.

Tolerances: VB is floating point with a different (field-parallel, tree-reduced) summation order than the
reference's sequential loops: north_star allows 1e-4 relative on test RMSE and free energy per iteration;
these tests hold it to 1e-7 on small inputs. MCMC uses a different RNG stream, so it is matched in
distribution (0.5 % on >= 1M rows; a looser bound on the small cases here) and exactly with sampling off."""
import numpy as np
import pytest

import oracle_binding as ob
import svbfm_b200 as sv
from helpers import make_learner, ragged, rel, synth, to_csc, two_field

pytestmark = pytest.mark.gpu

VB_TOL = 1e-7


def run_vb(tr, te, K, iters, **kw):
    orc = ob.Oracle("vb", tr, te, K=K, seed=42, k0=kw.get("k0", 1), k1=kw.get("k1", 1), groups=kw.get("groups"))
    L = make_learner("vb", tr, te, K, num_iter=iters, **kw)
    hist = L.learn(to_csc(tr), to_csc(te))
    for it, s in enumerate(hist):
        o = orc.iterate()
        assert rel(s.test_rmse, o.test_rmse) < VB_TOL, (it, s.test_rmse, o.test_rmse)
        assert rel(s.train_stat, o.train_stat) < VB_TOL, (it, s.train_stat, o.train_stat)
        assert s.has_free_energy == o.has_free_energy
        assert rel(s.free_energy, o.free_energy) < VB_TOL, (it, s.free_energy, o.free_energy)
        assert rel(s.alpha, o.alpha) < VB_TOL
    return L, orc


def test_vb_two_field_onehot(built):
    tr, te = two_field(20000, 2000, 300, 200)
    L, orc = run_vb(tr, te, K=4, iters=8)
    info = L.engine.info()
    assert info["num_runs"] == 2 and info["all_ones"] == 1 and info["uniform_row_nnz"] == 2 and info["rows_reordered"] == 1 and (info["fused_schedule"] & 1) == 1
    # state and residuals agree with the oracle's caches (caller case order)
    e_o, t_o = orc.get_train_cache()
    e = L.engine.get_residuals()
    assert np.max(np.abs(e - e_o)) < 1e-9
    assert rel(L.engine.get_sum_t(), t_o.sum()) < 1e-10
    so, sg = orc.get_state(), L.engine.get_state()
    for k in ("w_mean", "w_var", "v_mean", "v_var"):
        assert np.max(np.abs(so[k] - sg[k])) < 1e-9, k
    ho, hg = orc.get_hyper(), L.engine.get_hyper()
    assert np.allclose(ho["sigma_v"], hg["sigma_v"], rtol=1e-9) and np.allclose(ho["sigma_w"], hg["sigma_w"], rtol=1e-9)
    assert np.max(np.abs(L.engine.predict() - orc.get_test_pred())) < 1e-9
    assert L.engine.copies_max_diff() == 0.0     # the two residual copies of the stream schedule are bit-identical


def test_vb_values_no_reorder(built):
    tr, te = two_field(5000, 500, 100, 80, seed=5, values=True)
    L, _ = run_vb(tr, te, K=3, iters=5, flags=sv.FLAG_NO_ROW_REORDER)
    assert L.engine.info()["all_ones"] == 0


def test_vb_ragged_multihot(built):
    tr, te = ragged(3000, 400, 60)
    L, _ = run_vb(tr, te, K=3, iters=4)
    assert L.engine.info()["uniform_row_nnz"] == 0 and L.engine.info()["num_runs"] > 2


def test_vb_groups_and_small_tiles(built):
    tr, te = two_field(8000, 800, 150, 120, seed=9)
    D = max(tr.n_feat, te.n_feat) + 1
    groups = np.zeros(D, dtype=np.uint32)
    groups[150:] = 1
    L, _ = run_vb(tr, te, K=2, iters=4, groups=groups, tile_entries=32)   # tile_entries=32 forces heavy columns
    assert L.engine.info()["num_tiles"] > 270


@pytest.mark.parametrize("k0,k1", [(0, 1), (1, 0), (0, 0)])
def test_vb_k0_k1_switches(built, k0, k1):
    tr, te = two_field(4000, 400, 80, 60, seed=11)
    run_vb(tr, te, K=2, iters=3, k0=k0, k1=k1)


def test_vb_k_zero_and_empty_test_columns(built):
    tr, te = two_field(3000, 300, 50, 40, seed=13)
    run_vb(tr, te, K=0, iters=3)


def test_mcmc_als_exact(built):
    """do_sample=0 (-method als): no random numbers, so the sweep must match the oracle to rounding."""
    tr, te = two_field(20000, 2000, 300, 200, seed=21)
    orc = ob.Oracle("mcmc", tr, te, K=4, seed=42, do_sample=False, do_multilevel=False)
    L = make_learner("mcmc", tr, te, 4, num_iter=6, do_sample=False, do_multilevel=False)
    L.fm.reg0, L.fm.regw, L.fm.regv = 0.0, 0.0, 0.0
    hist = L.learn(to_csc(tr), to_csc(te))
    for it, s in enumerate(hist):
        o = orc.iterate()
        assert rel(s.test_rmse, o.test_rmse) < 1e-7, (it, s.test_rmse, o.test_rmse)
        assert rel(s.train_stat, o.train_stat) < 1e-7, (it, s.train_stat, o.train_stat)
        assert rel(s.rmse_this, o.rmse_this) < 1e-7


def test_mcmc_sampling_distribution(built):
    tr, te = two_field(200000, 20000, 1000, 800, seed=23)
    orc_vals = []
    for seed in (42, 43, 44):
        orc = ob.Oracle("mcmc", tr, te, K=4, seed=seed)
        for _ in range(15):
            o = orc.iterate()
        orc_vals.append(o.test_rmse)
    L = make_learner("mcmc", tr, te, 4, num_iter=15)
    hist = L.learn(to_csc(tr), to_csc(te))
    ref = float(np.mean(orc_vals))
    assert rel(hist[-1].test_rmse, ref) < 0.005 + 2 * (max(orc_vals) - min(orc_vals)) / ref, (hist[-1].test_rmse, orc_vals)
    assert np.isfinite(hist[-1].train_stat) and hist[-1].nan_inf_count == 0


def test_vb_online_vs_oracle(built):
    """vb_online: same libc shuffle stream, batches as case subsets of the resident matrix (no batch files).
    The oracle is run to completion first: both sides replay the process-global libc rand() stream."""
    tr, te = two_field(6000, 600, 100, 80, seed=31)
    want = []
    orc = ob.Oracle("vb_online", tr, te, K=3, seed=42, num_batch=5)
    for _ in range(4):
        s = orc.iterate()
        want.append((s.test_rmse, s.free_energy, s.alpha))
    so = orc.get_state()
    L = make_learner("vb_online", tr, te, 3, num_iter=4, num_batch=5)
    hist = L.learn(to_csc(tr), to_csc(te))
    for it, s in enumerate(hist):
        assert rel(s.test_rmse, want[it][0]) < VB_TOL, (it, s.test_rmse, want[it][0])
        assert s.has_free_energy and rel(s.free_energy, want[it][1]) < VB_TOL, (it, s.free_energy, want[it][1])
        assert rel(s.alpha, want[it][2]) < VB_TOL
    sg = L.engine.get_state()
    for k in ("w_mean", "w_var", "v_mean", "v_var"):
        assert np.max(np.abs(so[k] - sg[k])) < 1e-9, k


@pytest.mark.parametrize("values,k1,tile_entries", [(False, 1, 32), (True, 1, 64), (False, 0, 0), (True, 0, 256)])
def test_vb_online_stream_equals_general(built, monkeypatch, values, k1, tile_entries):
    """vb_online on two complete fields sweeps every batch with the stream schedule on the batch's own entries (per-epoch index
    lists); the masked general schedule (validated against the oracle above and below) must give the same statistics."""
    tr, te = two_field(9000, 900, 140, 110, seed=33, values=values)
    out = []
    for general in (False, True):
        if general:
            monkeypatch.setenv("SVBFM_NO_VBO_STREAM", "1")
        L = make_learner("vb_online", tr, te, 3, num_iter=3, num_batch=7, k1=k1, tile_entries=tile_entries)
        hist = L.learn(to_csc(tr), to_csc(te))
        assert (L.engine.info()["fused_schedule"] & 1) == (0 if general else 1)
        out.append(([(s.test_rmse, s.free_energy, s.alpha) for s in hist], L.engine.get_state()))
        L.engine.close()
        monkeypatch.delenv("SVBFM_NO_VBO_STREAM", raising=False)
    for a, b in zip(out[0][0], out[1][0]):
        assert all(rel(x, y) < 1e-9 for x, y in zip(a, b)), (a, b)
    for k in ("w_mean", "w_var", "v_mean", "v_var"):
        assert np.max(np.abs(out[0][1][k] - out[1][1][k])) < 1e-10, k


def test_vb_online_ragged_groups(built):
    tr, te = ragged(2500, 300, 40, seed=41)
    D = max(tr.n_feat, te.n_feat)
    groups = (np.arange(D) * 2 // D).astype(np.uint32)
    want = []
    orc = ob.Oracle("vb_online", tr, te, K=2, seed=7, num_batch=4, groups=groups)
    for _ in range(3):
        s = orc.iterate()
        want.append((s.test_rmse, s.free_energy))
    L = make_learner("vb_online", tr, te, 2, seed=7, num_iter=3, num_batch=4, groups=groups)
    hist = L.learn(to_csc(tr), to_csc(te))
    for it, s in enumerate(hist):
        assert rel(s.test_rmse, want[it][0]) < VB_TOL, (it, s.test_rmse, want[it][0])
        assert rel(s.free_energy, want[it][1]) < VB_TOL


def test_vb_block_cut_tiles(built, monkeypatch):
    """Big columns of gather runs are cut at case-block boundaries and executed block-major (L2 blocking):
    only the schedule changes, the sums per column keep their tile order."""
    monkeypatch.setenv("SVBFM_BLOCK_CASES", "1000")
    monkeypatch.setenv("SVBFM_NO_FUSE", "1")          # two-field data would otherwise take the stream schedule
    tr, te = two_field(20000, 2000, 300, 200, seed=61)
    L, _ = run_vb(tr, te, K=3, iters=4, tile_entries=64)
    assert L.engine.info()["num_tiles"] > 700 and L.engine.info()["fused_schedule"] == 0
    monkeypatch.delenv("SVBFM_NO_FUSE")
    monkeypatch.setenv("SVBFM_BLOCK_CASES", "700")
    tr, te = ragged(6000, 300, 30, seed=62)
    run_vb(tr, te, K=2, iters=3, tile_entries=32)


def test_stream_equals_general_schedule(built, monkeypatch):
    """The two-copy stream schedule (info: fused_schedule = 1) and the general per-run schedule are the same algorithm:
    identical statistics to rounding, on one-hot data and on two-field data with real values."""
    for values in (False, True):
        tr, te = two_field(15000, 1500, 250, 180, seed=71, values=values)
        out = []
        for nofuse in ("", "1"):
            if nofuse:
                monkeypatch.setenv("SVBFM_NO_FUSE", "1")
            else:
                monkeypatch.delenv("SVBFM_NO_FUSE", raising=False)
            L = make_learner("vb", tr, te, 3, num_iter=4)
            hist = L.learn(to_csc(tr), to_csc(te))
            assert (L.engine.info()["fused_schedule"] & 1) == (0 if nofuse else 1)
            out.append([(s.test_rmse, s.free_energy, s.train_stat) for s in hist])
            L.engine.close()
        for a, b in zip(*out):
            assert all(rel(x, y) < 1e-10 for x, y in zip(a, b)), (a, b)
        monkeypatch.delenv("SVBFM_NO_FUSE", raising=False)
        run_vb(tr, te, K=3, iters=4)


@pytest.mark.parametrize("tile_entries", [32, 64, 256])
def test_stream_schedule_small_tiles(built, tile_entries):
    """Stream schedule with tiny implicit tiles: columns inside one tile, columns crossing tile borders (light spans) and
    columns spanning many tiles (k_combine_span), empty columns in both fields, x != 1; both residual copies stay identical."""
    for values, method in ((False, "vb"), (True, "vb"), (False, "mcmc")):
        tr, te = two_field(30000, 3000, 400, 300, seed=91 + tile_entries, values=values)
        if method == "vb":
            L, _ = run_vb(tr, te, K=3, iters=4, tile_entries=tile_entries)
        else:
            orc = ob.Oracle("mcmc", tr, te, K=3, seed=42, do_sample=False, do_multilevel=False)
            L = make_learner("mcmc", tr, te, 3, num_iter=4, do_sample=False, do_multilevel=False, tile_entries=tile_entries)
            L.fm.reg0, L.fm.regw, L.fm.regv = 0.0, 0.0, 0.0
            for s in L.learn(to_csc(tr), to_csc(te)):
                o = orc.iterate()
                assert rel(s.test_rmse, o.test_rmse) < 1e-7 and rel(s.train_stat, o.train_stat) < 1e-7
        info = L.engine.info()
        assert (info["fused_schedule"] & 1) == 1 and info["num_tiles"] >= 2 * (30000 // tile_entries)
        assert L.engine.copies_max_diff() == 0.0
        L.engine.close()


def test_stream_schedule_id_gaps(built):
    """Long stretches of never-used feature ids inside both fields (more than a 32-column window, so the column walk of
    k_stream has to skip whole windows by binary search), for vb, als and vb_online, with small tiles."""
    r = np.random.default_rng(5)
    U, I, N, Nt = 400, 300, 8000, 800
    users = np.concatenate([np.arange(0, 30), np.arange(130, 160), np.arange(395, 400)])       # gaps of 100 and 235 ids
    items = np.concatenate([np.arange(2, 12), np.arange(120, 170), np.arange(290, 300)])       # gaps of 108 and 120 ids

    def make(n, seed):
        rr = np.random.default_rng(seed)
        u, i = rr.choice(users, n), rr.choice(items, n)
        y = np.clip(np.round(3.5 + 0.01 * (u % 7) - 0.02 * (i % 5) + rr.normal(0, 0.8, n)), 1, 5).astype(np.float32)
        return ob.Csr(*synth.to_csr(u.astype(np.uint32), i.astype(np.uint32), y, U))
    tr, te = make(N, 1), make(Nt, 2)
    for te_ in (32, 256):
        L, _ = run_vb(tr, te, K=3, iters=3, tile_entries=te_)
        assert (L.engine.info()["fused_schedule"] & 1) == 1 and L.engine.copies_max_diff() == 0.0
        L.engine.close()
    orc = ob.Oracle("mcmc", tr, te, K=2, seed=42, do_sample=False, do_multilevel=False)
    L = make_learner("mcmc", tr, te, 2, num_iter=3, do_sample=False, do_multilevel=False, tile_entries=64)
    L.fm.reg0, L.fm.regw, L.fm.regv = 0.0, 0.0, 0.0
    for s in L.learn(to_csc(tr), to_csc(te)):
        o = orc.iterate()
        assert rel(s.test_rmse, o.test_rmse) < 1e-7 and rel(s.train_stat, o.train_stat) < 1e-7
    L.engine.close()
    want = []
    orc = ob.Oracle("vb_online", tr, te, K=2, seed=42, num_batch=6)
    for _ in range(3):
        s = orc.iterate()
        want.append((s.test_rmse, s.free_energy))
    L = make_learner("vb_online", tr, te, 2, num_iter=3, num_batch=6, tile_entries=32)
    for it, s in enumerate(L.learn(to_csc(tr), to_csc(te))):
        assert rel(s.test_rmse, want[it][0]) < VB_TOL and rel(s.free_energy, want[it][1]) < VB_TOL
    assert (L.engine.info()["fused_schedule"] & 1) == 1


def test_stream_schedule_sorted_input(built, monkeypatch):
    """Cases already sorted by the first field (no re-ordering needed once the rank layout is off) still take the stream schedule."""
    monkeypatch.setenv("SVBFM_REC_RANK", "0")
    tr, te = two_field(12000, 1200, 200, 150, seed=95)
    order = np.argsort(tr.col[0::2], kind="stable")
    idx = np.stack([2 * order, 2 * order + 1], axis=1).ravel()
    tr2 = ob.Csr(tr.rowptr.copy(), tr.col[idx].copy(), tr.val[idx].copy(), tr.y[order].copy())
    L, _ = run_vb(tr2, te, K=3, iters=3)
    info = L.engine.info()
    assert (info["fused_schedule"] & 1) == 1 and info["rows_reordered"] == 0
    assert L.engine.copies_max_diff() == 0.0


@pytest.mark.parametrize("K", [40, 70, 130])
def test_two_field_train_prediction_wide_k(built, K, monkeypatch):
    """k_predict2 (transposed parameters, warp per case) with 2, 4 and 8 factor slots per lane, one-hot and real values, vb and
    mcmc/als; the residuals it produces equal those of the general case-wise kernel (k_predict) to rounding."""
    for values in (False, True):
        tr, te = two_field(6000, 600, 90, 70, seed=97, values=values)
        L, _ = run_vb(tr, te, K=K, iters=2)
        e_fast = L.engine.get_residuals()
        L.engine.close()
        monkeypatch.setenv("SVBFM_NO_PREDICT2", "1")
        L2 = make_learner("vb", tr, te, K, num_iter=2)
        L2.learn(to_csc(tr), to_csc(te))
        assert np.max(np.abs(L2.engine.get_residuals() - e_fast)) < 1e-9
        L2.engine.close()
        monkeypatch.delenv("SVBFM_NO_PREDICT2")
    tr, te = two_field(6000, 600, 90, 70, seed=98)
    orc = ob.Oracle("mcmc", tr, te, K=K, seed=42, do_sample=False, do_multilevel=False)
    L = make_learner("mcmc", tr, te, K, num_iter=3, do_sample=False, do_multilevel=False)
    L.fm.reg0, L.fm.regw, L.fm.regv = 0.0, 0.0, 0.0
    for s in L.learn(to_csc(tr), to_csc(te)):
        o = orc.iterate()
        assert rel(s.test_rmse, o.test_rmse) < 1e-7 and rel(s.train_stat, o.train_stat) < 1e-7


def test_reset_reuses_handle(built):
    """svbfm_reset: a second learn() on the same handle (same data, same initial state) repeats the first one."""
    tr, te = two_field(10000, 1000, 200, 150, seed=81)
    L = make_learner("vb", tr, te, 3, num_iter=3)
    h1 = [(s.test_rmse, s.free_energy) for s in L.learn(to_csc(tr), to_csc(te))]
    E = L.engine
    E.reset()
    E.set_csc(sv.TRAIN, to_csc(tr)); E.set_csc(sv.TEST, to_csc(te)); E.set_state(L._state); E.begin()
    h2 = [(s.test_rmse, s.free_energy) for s in E.run(3)]
    assert h1 == h2


def test_vb_online_long_spans_in_a_batch(built):
    """vb_online on the stream schedule with 32-entry tiles and ten users: every user's column spans ~60 tiles of a batch, which the
    lanes of a warp sum together in k_finalize_vbo (a single thread used to walk them: the critical path of every finalize at 200 M)."""
    tr, te = two_field(40000, 500, 10, 700, seed=77)
    want = []
    orc = ob.Oracle("vb_online", tr, te, K=2, seed=42, num_batch=2)
    for _ in range(2):
        s = orc.iterate()
        want.append((s.test_rmse, s.free_energy, s.alpha))
    L = make_learner("vb_online", tr, te, 2, num_iter=2, num_batch=2, tile_entries=32)
    for it, s in enumerate(L.learn(to_csc(tr), to_csc(te))):
        assert rel(s.test_rmse, want[it][0]) < VB_TOL and rel(s.free_energy, want[it][1]) < VB_TOL and rel(s.alpha, want[it][2]) < VB_TOL
    assert L.engine.info()["fused_schedule"] & 1
    L.engine.close()


def test_values_may_be_null_for_one_hot_data(built):
    """x = NULL in svbfm_set_csc / svbfm_set_csr says "every value is 1": same statistics, bit for bit, as an explicit array of ones,
    column-wise and row-wise (the device transpose)."""
    tr, te = two_field(6000, 600, 100, 80, seed=5)
    D = max(tr.n_feat, te.n_feat) + 1
    out = []
    for mode in ("ones", "null", "rows_null"):
        E = sv.Engine("vb", D, 3, 1, 1, float(tr.y.min()), float(tr.y.max()), seed=42)
        a, b = to_csc(tr), to_csc(te)
        if mode == "null":
            a.x = b.x = None
        if mode == "rows_null":
            E.set_csr(sv.TRAIN, tr.rowptr, tr.col, None, tr.y, tr.n_feat); E.set_csr(sv.TEST, te.rowptr, te.col, None, te.y, te.n_feat)
        else:
            E.set_csc(sv.TRAIN, a); E.set_csc(sv.TEST, b)
        assert E.info()["all_ones"] == 1 and E.info()["fused_schedule"] & 1
        E.set_state(sv.host_init_state(42, D, 3, 0.1, sv.VB)); E.begin()
        out.append([(s.test_rmse, s.free_energy, s.alpha) for s in E.run(3)])
        E.close()
    assert out[0] == out[1] == out[2]


def test_reset_onto_a_larger_split_vb_online(built):
    """A long-lived vb_online handle that is reset onto a LARGER train split: the per-case / per-entry batch buffers follow the new
    split (they used to keep the first split's size). Same statistics as a fresh handle on the larger split, bit for bit."""
    small, te_s = two_field(3000, 300, 120, 90, seed=83)
    big, te = two_field(12000, 800, 120, 90, seed=84)
    D = max(big.n_feat, te.n_feat, small.n_feat, te_s.n_feat)
    nb = 4

    def epochs(E, tr, k):
        batch = (np.random.default_rng(5).permutation(tr.n_rows) % nb).astype(np.uint32)
        return [(s.test_rmse, s.free_energy) for s in (E.vb_online_epoch(batch, nb) for _ in range(k))]

    state = sv.host_init_state(42, D, 3, 0.1, sv.VB_ONLINE)
    E = sv.Engine("vb_online", D, 3, 1, 1, 1.0, 5.0, seed=42)
    E.set_csc(sv.TRAIN, to_csc(small)); E.set_csc(sv.TEST, to_csc(te_s)); E.set_state(state); E.begin()
    epochs(E, small, 1)
    E.reset()
    E.set_csc(sv.TRAIN, to_csc(big)); E.set_csc(sv.TEST, to_csc(te)); E.set_state(state); E.begin()
    again = epochs(E, big, 2)
    E.close()
    F = sv.Engine("vb_online", D, 3, 1, 1, 1.0, 5.0, seed=42)
    F.set_csc(sv.TRAIN, to_csc(big)); F.set_csc(sv.TEST, to_csc(te)); F.set_state(state); F.begin()
    fresh = epochs(F, big, 2)
    F.close()
    assert again == fresh


@pytest.mark.parametrize("method", ["vb", "mcmc", "vb_online"])
def test_degenerate_splits(built, method):
    """An empty test split (the reference divides by zero cases and prints `Test=nan`, vbs.h:261-279 / mcmcs.h:226-233), a single
    train case (one tile, one entry per column, every other column empty) and a train split that is smaller than one warp."""
    kw = dict(do_sample=False, do_multilevel=False) if method == "mcmc" else dict(num_batch=2) if method == "vb_online" else {}
    tr, te = two_field(500, 50, 20, 15, seed=5)
    empty = ob.Csr(np.zeros(1, dtype=np.uint64), np.zeros(0, dtype=np.uint32), np.zeros(0, dtype=np.float32), np.zeros(0, dtype=np.float32))
    one, _ = two_field(1, 5, 20, 15, seed=6)
    few, _ = two_field(7, 5, 20, 15, seed=7)
    cases = [(tr, empty), (few, te)] + ([(one, te)] if method != "vb_online" else [])      # vb_online: one case in two batches leaves a batch empty
    for a, b in cases:
        orc = ob.Oracle(method, a, b, K=2, seed=42, **kw)
        want = [orc.iterate() for _ in range(3)]          # the oracle first: vb_online replays the process-global libc rand() stream on both sides
        L = make_learner(method, a, b, 2, num_iter=3, **kw)
        same = lambda x, y: (np.isnan(x) and np.isnan(y)) or (abs(x) < 1e-8 and abs(y) < 1e-8) or rel(x, y) < VB_TOL     # an exact fit leaves rounding noise
        for it, (s, o) in enumerate(zip(L.learn(to_csc(a), to_csc(b)), want)):
            assert np.isnan(o.test_rmse) == (b.n_rows == 0)
            assert same(s.test_rmse, o.test_rmse), (method, a.n_rows, it, s.test_rmse, o.test_rmse)
            assert same(s.train_stat, o.train_stat), (method, a.n_rows, it, s.train_stat, o.train_stat)
            if method != "mcmc":
                assert same(s.free_energy, o.free_energy), (method, a.n_rows, it, s.free_energy, o.free_energy)
        L.engine.close()
