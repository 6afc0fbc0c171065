"""k_stream with its all-ones streams staged through a shared-memory ring of bulk copies (cp.async.bulk + mbarrier; the default since
round 2, first run on a B200 in gpurun call r2a) against the plain-load variant (SVBFM_STREAM_TMA=0): bit-identical. A wrong
barrier phase would spin forever, so the checks run in a child process under a timeout: a hang or a sticky CUDA error cannot
take the rest of the GPU suite with it."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if __name__ == "__main__":       # the child process: same import roots as tests/conftest.py sets up
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]

import numpy as np
import pytest

import oracle_binding as ob
from helpers import make_learner, rel, to_csc, two_field

pytestmark = pytest.mark.gpu
VB_TOL = 1e-7
CASES = [(0, 20000), (64, 20000), (128, 20000), (256, 20002), (512, 20002), (1024, 30001)]


def _emulated():
    return "emu" in os.path.basename(os.environ.get("SVBFM_LIB", ""))


@pytest.mark.parametrize("tile_entries,n", CASES)
def test_stream_tma_ring(built, tile_entries, n):
    p = subprocess.run([sys.executable, os.path.abspath(__file__), str(tile_entries), str(n)], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "TMA_RING_OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]


def check_stream_tma_ring(tile_entries, n):
    """The streams of k_stream staged through a per-warp shared-memory ring of bulk copies. Same
    arithmetic in the same order as the plain kernel: identical statistics (bit for bit) and identical residual copies.
    n = 20002 / 30001: the second field's streams do not start on a 16-byte boundary (plain kernel for that side), partial
    last batches and tiles."""
    tr, te = two_field(n, 2000, 300, 200, seed=31)
    out = []
    for tma in ("0", "1"):
        os.environ["SVBFM_STREAM_TMA"] = tma
        L = make_learner("vb", tr, te, 3, num_iter=3, tile_entries=tile_entries)
        out.append([(s.test_rmse, s.free_energy, s.alpha, s.train_stat) for s in L.learn(to_csc(tr), to_csc(te))])
        assert L.engine.info()["fused_schedule"] & 5 == (5 if tma == "1" else 1)
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


if __name__ == "__main__":
    check_stream_tma_ring(int(sys.argv[1]), int(sys.argv[2]))
    print("TMA_RING_OK")
