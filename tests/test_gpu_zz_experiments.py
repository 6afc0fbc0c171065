"""Opt-in variants of the hot path (environment knobs, off by default) against the oracle. Kept in a file that sorts last:
these paths are experiments for the next measurement round, the default path's parity tests come first."""
import numpy as np
import pytest

import oracle_binding as ob
from helpers import make_learner, rel, to_csc, two_field

pytestmark = pytest.mark.gpu
VB_TOL = 1e-7


@pytest.mark.parametrize("tile_entries", [0, 64])
def test_rec_rank_layout(built, monkeypatch, tile_entries):
    """SVBFM_REC_RANK=1: second-field records by popularity rank, cases of a first-field column ordered by that rank.
    Same algorithm, another case order: statistics, parameters and residuals (caller order) equal the oracle's."""
    monkeypatch.setenv("SVBFM_REC_RANK", "1")
    for values in (False, True):
        tr, te = two_field(20000, 2000, 300, 200, seed=11, values=values)
        orc = ob.Oracle("vb", tr, te, K=3, seed=42)
        L = make_learner("vb", tr, te, 3, num_iter=4, tile_entries=tile_entries)
        for it, s in enumerate(L.learn(to_csc(tr), to_csc(te))):
            o = orc.iterate()
            assert rel(s.test_rmse, o.test_rmse) < VB_TOL and rel(s.free_energy, o.free_energy) < VB_TOL and rel(s.alpha, o.alpha) < VB_TOL, it
        assert L.engine.info()["fused_schedule"] == 3
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
    assert L.engine.info()["fused_schedule"] == 3
    L.engine.close()
    orc = ob.Oracle("vb_online", tr, te, K=2, seed=42, num_batch=5)
    want = [orc.iterate() for _ in range(2)]
    L = make_learner("vb_online", tr, te, 2, num_iter=2, num_batch=5, tile_entries=tile_entries)
    for it, s in enumerate(L.learn(to_csc(tr), to_csc(te))):
        assert rel(s.test_rmse, want[it].test_rmse) < VB_TOL and rel(s.free_energy, want[it].free_energy) < VB_TOL
    assert L.engine.info()["fused_schedule"] == 3
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
        assert L.engine.info()["fused_schedule"] == 1
        L.engine.close()
    for a, b in zip(*out):
        assert all(rel(x, y) < 1e-10 for x, y in zip(a, b)), (a, b)


@pytest.mark.parametrize("tile_entries,n", [(0, 20000), (64, 20000), (256, 20002), (1024, 30001)])
def test_stream_tma_ring(built, monkeypatch, tile_entries, n):
    """SVBFM_STREAM_TMA=1: the streams of k_stream staged through a per-warp shared-memory ring of bulk copies. Same
    arithmetic in the same order as the plain kernel: identical statistics (bit for bit) and identical residual copies.
    n = 20002 / 30001: the second field's streams do not start on a 16-byte boundary (plain kernel for that side), partial
    last batches and tiles."""
    tr, te = two_field(n, 2000, 300, 200, seed=31)
    out = []
    for tma in ("0", "1"):
        monkeypatch.setenv("SVBFM_STREAM_TMA", tma)
        L = make_learner("vb", tr, te, 3, num_iter=3, tile_entries=tile_entries)
        out.append([(s.test_rmse, s.free_energy, s.alpha, s.train_stat) for s in L.learn(to_csc(tr), to_csc(te))])
        assert L.engine.info()["fused_schedule"] == (5 if tma == "1" else 1)
        assert L.engine.copies_max_diff() == 0.0
        out.append(L.engine.get_residuals())
        L.engine.close()
    assert out[0] == out[2]
    assert np.array_equal(out[1], out[3])
    orc = ob.Oracle("vb", tr, te, K=3, seed=42)
    for a in out[2]:
        o = orc.iterate()
        assert rel(a[0], o.test_rmse) < VB_TOL and rel(a[1], o.free_energy) < VB_TOL
    # als through the same ring
    orc = ob.Oracle("mcmc", tr, te, K=2, seed=42, do_sample=False, do_multilevel=False)
    L = make_learner("mcmc", tr, te, 2, num_iter=3, do_sample=False, do_multilevel=False, tile_entries=tile_entries)
    L.fm.reg0, L.fm.regw, L.fm.regv = 0.0, 0.0, 0.0
    for s in L.learn(to_csc(tr), to_csc(te)):
        o = orc.iterate()
        assert rel(s.test_rmse, o.test_rmse) < VB_TOL and rel(s.train_stat, o.train_stat) < VB_TOL
    assert L.engine.copies_max_diff() == 0.0
    L.engine.close()
