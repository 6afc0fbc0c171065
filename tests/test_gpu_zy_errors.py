"""-m gpu: the error convention of the C-ABI (negative status + svbfm_last_error; the shells turn it into the reference's
`throw std::string`, libfm.cpp:521-527): bad arguments and bad data are rejected with a message and leave the handle usable."""
import numpy as np
import pytest

import svbfm_b200 as sv
from helpers import make_learner, to_csc, two_field

pytestmark = pytest.mark.gpu


def _csc(colptr, case_id, y, x=None, n=None, ncols=None):
    colptr = np.asarray(colptr, dtype=np.uint64)
    case_id = np.asarray(case_id, dtype=np.uint32)
    x = np.ones(len(case_id), dtype=np.float32) if x is None else np.asarray(x, dtype=np.float32)
    d = sv.CscData(colptr, case_id, x, np.asarray(y, dtype=np.float32))
    if n is not None:
        d.num_cases = n
    if ncols is not None:
        d.num_feature = ncols
    return d


def test_bad_arguments_are_rejected(built):
    with pytest.raises(sv.SvbfmError, match="unknown method"):
        sv.Engine(7, 10, 2)
    with pytest.raises(sv.SvbfmError, match="bad dimensions"):
        sv.Engine("vb", 0, 2)
    with pytest.raises(sv.SvbfmError, match="task"):
        sv.Engine("vb", 10, 2, task=1)            # classification is on the mcmc path only
    E = sv.Engine("vb", 6, 2, 1, 1, 1.0, 5.0)
    good = _csc([0, 2, 3, 3, 5, 6], [0, 2, 1, 0, 1, 2], [3, 4, 5])      # 3 cases, columns 0..4
    with pytest.raises(sv.SvbfmError, match="num_cols exceeds"):
        E.set_csc(sv.TRAIN, _csc([0] * 8, [], [1.0], ncols=7))
    with pytest.raises(sv.SvbfmError, match="not monotone"):
        E.set_csc(sv.TRAIN, _csc([0, 2, 1, 3, 5, 6], [0, 2, 1, 0, 1, 2], [3, 4, 5]))
    with pytest.raises(sv.SvbfmError, match="colptr\\[0\\]"):
        E.set_csc(sv.TRAIN, _csc([1, 2, 3, 3, 5, 6], [0, 2, 1, 0, 1, 2], [3, 4, 5]))
    with pytest.raises(sv.SvbfmError, match="case id out of range"):
        E.set_csc(sv.TRAIN, _csc([0, 2, 3, 3, 5, 6], [0, 2, 1, 0, 1, 9], [3, 4, 5]))
    with pytest.raises(sv.SvbfmError, match="occurs twice"):
        E.set_csc(sv.TRAIN, _csc([0, 2, 3, 3, 5, 6], [0, 0, 1, 0, 1, 2], [3, 4, 5]))     # case 0 twice in column 0
    with pytest.raises(sv.SvbfmError, match="set_state must be called first"):
        E.begin()
    # the handle is still usable after all of that
    E.set_csc(sv.TRAIN, good)
    E.set_csc(sv.TEST, good)
    E.set_state(sv.host_init_state(42, 6, 2, 0.1, sv.VB))
    E.begin()
    with pytest.raises(sv.SvbfmError, match="after svbfm_begin"):
        E.set_csc(sv.TRAIN, good)
    with pytest.raises(sv.SvbfmError, match="non-mcmc handle"):
        E._ck(sv.lib().svbfm_mcmc_sweep(E.h, None), "svbfm_mcmc_sweep")
    s = E.run(2)
    assert np.isfinite(s[-1].test_rmse) and np.isfinite(s[-1].free_energy)
    E.close()


def test_vb_online_batch_ids_are_checked(built):
    tr, te = two_field(2000, 200, 40, 30, seed=3)
    L = make_learner("vb_online", tr, te, 2, num_iter=1, num_batch=4)
    L._make_engine()
    E = L.engine
    E.set_csc(sv.TRAIN, to_csc(tr))
    E.set_csc(sv.TEST, to_csc(te))
    E.set_state(L._state)
    with pytest.raises(sv.SvbfmError, match="svbfm_begin must be called first"):
        E.vb_online_epoch(np.zeros(2000, dtype=np.uint32), 4)
    E.begin()
    with pytest.raises(sv.SvbfmError, match="batch id out of range"):
        E.vb_online_epoch(np.full(2000, 4, dtype=np.uint32), 4)
    s = E.vb_online_epoch((np.arange(2000) % 4).astype(np.uint32), 4)
    assert np.isfinite(s.test_rmse)
    E.close()
