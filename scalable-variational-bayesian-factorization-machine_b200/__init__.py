"""B200-native VB / vb_online / MCMC coordinate sweep for factorization machines.

Python mirror of the reference's learner interface (`class fm_learn`, reference
src/libfm/src/fm_learn.h:38-265, and its subclasses fm_learn_vb_simultaneous / fm_learn_mcmc_simultaneous /
fm_learn_vb_online_simultaneous) on top of the C-ABI in include/svbfm.h. All compute happens in
libsvbfm.so (hand-written sm_100a CUDA); there is no CPU or PyTorch fallback -- if the library or a B200
is missing the calls raise.

The C++ host side (host/: libFM-compatible CLI, convert, transpose) binds the same C-ABI.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SVBFM_LIB") or os.path.join(_HERE, "libsvbfm.so")   # SVBFM_LIB: another build of the same ABI

VB, VB_ONLINE, MCMC = 0, 1, 2
TRAIN, TEST, TRAIN_SECOND = 0, 1, 2     # TRAIN_SECOND: the second residual copy's shard (several GPUs; include/svbfm.h)
METHODS = {"vb": VB, "vb_online": VB_ONLINE, "mcmc": MCMC}
FLAG_NO_ROW_REORDER = 1
FLAG_MCMC_NO_REPREDICT = 2
COMM_ID_BYTES = 128


class SvbfmError(RuntimeError):
    """Raised where the reference would `throw std::string` (libfm.cpp:521-525)."""


class Config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("method", C.c_int32), ("num_attribute", C.c_uint32),
                ("num_factor", C.c_int32), ("k0", C.c_int32), ("k1", C.c_int32), ("task", C.c_int32),
                ("min_target", C.c_double), ("max_target", C.c_double), ("device", C.c_int32),
                ("do_sample", C.c_int32), ("do_multilevel", C.c_int32), ("seed", C.c_uint64),
                ("reg0", C.c_double), ("regw", C.c_double), ("regv", C.c_double),
                ("tile_entries", C.c_uint32), ("flags", C.c_uint32)]


class IterStats(C.Structure):
    _fields_ = [("test_rmse", C.c_double), ("train_stat", C.c_double), ("free_energy", C.c_double),
                ("alpha", C.c_double), ("rmse_this", C.c_double), ("has_free_energy", C.c_int32),
                ("nan_inf_count", C.c_uint32), ("sweep_ms", C.c_float), ("predict_ms", C.c_float),
                ("free_energy_first", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Info(C.Structure):
    _fields_ = [("num_runs", C.c_uint32), ("num_tiles", C.c_uint32), ("uniform_row_nnz", C.c_uint32),
                ("all_ones", C.c_uint32), ("kernel_launches", C.c_uint64), ("device_bytes", C.c_uint64),
                ("train_nnz", C.c_uint64), ("rows_reordered", C.c_uint32), ("world_size", C.c_uint32),
                ("fused_schedule", C.c_uint32), ("exclusive_blocks", C.c_uint32)]


# every symbol include/svbfm.h declares (tests/test_abi.py checks the library exports exactly these)
ABI_SYMBOLS = [
    "svbfm_create", "svbfm_destroy", "svbfm_last_error", "svbfm_abi_version", "svbfm_comm_get_unique_id",
    "svbfm_comm_init", "svbfm_set_groups", "svbfm_set_csc", "svbfm_set_csr", "svbfm_transpose_csr", "svbfm_set_state", "svbfm_get_state",
    "svbfm_get_hyper", "svbfm_set_hyper", "svbfm_begin", "svbfm_vb_sweep", "svbfm_mcmc_sweep",
    "svbfm_vb_online_epoch", "svbfm_run", "svbfm_reset", "svbfm_predict", "svbfm_get_residuals", "svbfm_get_sum_t", "svbfm_set_residuals", "svbfm_copies_max_diff",
    "svbfm_get_info", "svbfm_set_stream", "svbfm_set_profile", "svbfm_get_profile", "svbfm_host_init_state", "svbfm_host_random_shuffle",
]

_lib = None


def lib():
    """Load libsvbfm.so; fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SvbfmError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        vp, u32p = C.c_void_p, C.c_void_p
        L.svbfm_create.argtypes = [C.POINTER(vp), C.POINTER(Config)]
        L.svbfm_destroy.argtypes = [vp]
        L.svbfm_last_error.argtypes = [vp]
        L.svbfm_last_error.restype = C.c_char_p
        L.svbfm_comm_get_unique_id.argtypes = [vp]
        L.svbfm_comm_init.argtypes = [vp, vp, C.c_int32, C.c_int32]
        L.svbfm_set_groups.argtypes = [vp, u32p, C.c_uint32]
        L.svbfm_set_csc.argtypes = [vp, C.c_int32, C.c_uint32, C.c_uint32, vp, vp, vp, vp]
        L.svbfm_set_csr.argtypes = [vp, C.c_int32, C.c_uint32, C.c_uint32, vp, vp, vp, vp]
        L.svbfm_transpose_csr.argtypes = [C.c_int32, C.c_uint32, C.c_uint32, vp, vp, vp, vp, vp, vp]
        L.svbfm_set_state.argtypes = [vp, C.c_double, C.c_double, vp, vp, vp, vp]
        L.svbfm_get_state.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), vp, vp, vp, vp]
        L.svbfm_get_hyper.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), vp, vp]
        L.svbfm_set_hyper.argtypes = [vp, C.c_double, C.c_double, vp, vp]
        L.svbfm_begin.argtypes = [vp]
        L.svbfm_vb_sweep.argtypes = [vp, C.POINTER(IterStats)]
        L.svbfm_mcmc_sweep.argtypes = [vp, C.POINTER(IterStats)]
        L.svbfm_vb_online_epoch.argtypes = [vp, vp, C.c_uint32, C.POINTER(IterStats)]
        L.svbfm_run.argtypes = [vp, C.c_uint32, C.POINTER(IterStats)]
        L.svbfm_reset.argtypes = [vp]
        L.svbfm_predict.argtypes = [vp, C.c_int32, vp]
        L.svbfm_get_residuals.argtypes = [vp, vp]
        L.svbfm_get_sum_t.argtypes = [vp, C.POINTER(C.c_double)]
        L.svbfm_set_residuals.argtypes = [vp, vp, C.c_double]
        L.svbfm_copies_max_diff.argtypes = [vp, C.POINTER(C.c_double)]
        L.svbfm_get_info.argtypes = [vp, C.POINTER(Info)]
        L.svbfm_set_stream.argtypes = [vp, vp]
        L.svbfm_set_profile.argtypes = [vp, C.c_int32]
        L.svbfm_get_profile.argtypes = [vp, vp, vp]
        L.svbfm_host_init_state.argtypes = [C.c_long, C.c_uint32, C.c_int32, C.c_double, C.c_int32,
                                            C.POINTER(C.c_double), C.POINTER(C.c_double), vp, vp, vp, vp]
        L.svbfm_host_random_shuffle.argtypes = [vp, C.c_uint32]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def transpose_csr(rowptr, feature_id, x, num_feature, device=0):
    """Device transpose CSR -> CSC (colptr uint64[num_feature + 1], case_id uint32[nnz], x float32[nnz]): the .xt of a .x."""
    rowptr = np.ascontiguousarray(rowptr, dtype=np.uint64); feature_id = np.ascontiguousarray(feature_id, dtype=np.uint32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    n, nnz = len(rowptr) - 1, int(rowptr[-1])
    colptr, case_id, xt = np.zeros(int(num_feature) + 1, dtype=np.uint64), np.zeros(nnz, dtype=np.uint32), np.zeros(nnz, dtype=np.float32)
    rc = lib().svbfm_transpose_csr(int(device), n, int(num_feature), _p(rowptr), _p(feature_id), _p(x), _p(colptr), _p(case_id), _p(xt))
    if rc != 0:
        msg = lib().svbfm_last_error(None)
        raise SvbfmError(f"svbfm_transpose_csr failed ({rc}): {msg.decode() if msg else ''}")
    return colptr, case_id, xt


def host_init_state(seed, D, K, init_stdev=0.1, method=VB):
    """Replay of the reference's libc-rand() initial state (host/init_state.h)."""
    w0m, w0v = C.c_double(), C.c_double()
    wm, wv = np.zeros(D), np.zeros(D)
    vm, vv = np.zeros((K, D)), np.zeros((K, D))
    lib().svbfm_host_init_state(int(seed), int(D), int(K), float(init_stdev), int(method), C.byref(w0m), C.byref(w0v),
                                _p(wm), _p(wv), _p(vm), _p(vv))
    return dict(w0_mean=w0m.value, w0_var=w0v.value, w_mean=wm, w_var=wv, v_mean=vm, v_var=vv)


class CscData:
    """`DataSubset` as the learners see it: data_t (CSC of X, rows = features) + target (Data.h:87-89)."""

    def __init__(self, colptr, case_id, x, target, num_cases=None):
        self.colptr = np.ascontiguousarray(colptr, dtype=np.uint64)
        self.case_id = np.ascontiguousarray(case_id, dtype=np.uint32)
        self.x = None if x is None else np.ascontiguousarray(x, dtype=np.float32)      # None: every value is 1 (not shipped)
        self.target = np.ascontiguousarray(target, dtype=np.float32)
        self.num_cases = int(num_cases) if num_cases is not None else len(self.target)
        self.num_feature = len(self.colptr) - 1          # data_t->getNumRows() = max id + 1
        self.min_target = float(self.target.min()) if len(self.target) else float(np.finfo(np.float32).max)
        self.max_target = float(self.target.max()) if len(self.target) else -float(np.finfo(np.float32).max)

    @staticmethod
    def from_csr(rowptr, col, val, y, n_feat=None):
        """Counting transpose in case order (what Data::create_data_t does, Data.h:457-509), vectorised."""
        rowptr = np.asarray(rowptr, dtype=np.int64)
        col = np.asarray(col, dtype=np.int64)
        n = len(rowptr) - 1
        n_feat = int(n_feat) if n_feat is not None else (int(col.max()) + 1 if len(col) else 0)
        rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(rowptr))
        order = np.argsort(col, kind="stable")
        colptr = np.zeros(n_feat + 1, dtype=np.uint64)
        np.cumsum(np.bincount(col, minlength=n_feat), out=colptr[1:])
        return CscData(colptr, rows[order].astype(np.uint32), np.asarray(val, dtype=np.float32)[order], y)


class Engine:
    """Thin RAII wrapper of one C-ABI handle."""

    def __init__(self, method, num_attribute, num_factor, k0=1, k1=1, min_target=0.0, max_target=0.0, device=0,
                 seed=42, do_sample=True, do_multilevel=True, reg=(0.0, 0.0, 0.0), tile_entries=0, flags=0, task=0):
        self.method = METHODS[method] if isinstance(method, str) else int(method)
        cfg = Config(C.sizeof(Config), self.method, int(num_attribute), int(num_factor), int(bool(k0)), int(bool(k1)), int(task),
                     float(min_target), float(max_target), int(device), int(do_sample), int(do_multilevel), int(seed),
                     float(reg[0]), float(reg[1]), float(reg[2]), int(tile_entries), int(flags))
        self.D, self.K, self.G = int(num_attribute), int(num_factor), 1
        self.h = C.c_void_p()
        rc = lib().svbfm_create(C.byref(self.h), C.byref(cfg))
        if rc != 0:
            msg = lib().svbfm_last_error(None)
            self.h = None
            raise SvbfmError(f"svbfm_create failed ({rc}): {msg.decode() if msg else ''}")
        self.n_train = self.n_test = 0

    def _ck(self, rc, what):
        if rc != 0:
            msg = lib().svbfm_last_error(self.h)
            raise SvbfmError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def comm_init(self, unique_id, rank, world_size):
        buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(bytes(unique_id))
        self._ck(lib().svbfm_comm_init(self.h, buf, rank, world_size), "svbfm_comm_init")

    def set_groups(self, attr_group):
        g = np.ascontiguousarray(attr_group, dtype=np.uint32)
        assert len(g) == self.D
        self.G = int(g.max()) + 1
        self._ck(lib().svbfm_set_groups(self.h, _p(g), self.G), "svbfm_set_groups")

    def set_csc(self, split, data):
        self._ck(lib().svbfm_set_csc(self.h, split, data.num_cases, data.num_feature, _p(data.colptr), _p(data.case_id),
                                     _p(data.x), _p(data.target)), "svbfm_set_csc")
        if split == TRAIN:
            self.n_train = data.num_cases
        elif split == TEST:
            self.n_test = data.num_cases

    def set_csr(self, split, rowptr, feature_id, x, target, num_feature):
        """The split row-wise (CSR of the cases, like a parsed text file or a binary .x file): the device builds the transposed matrix."""
        rowptr = np.ascontiguousarray(rowptr, dtype=np.uint64); feature_id = np.ascontiguousarray(feature_id, dtype=np.uint32)
        x = None if x is None else np.ascontiguousarray(x, dtype=np.float32)
        target = np.ascontiguousarray(target, dtype=np.float32)
        n = len(rowptr) - 1
        self._ck(lib().svbfm_set_csr(self.h, split, n, int(num_feature), _p(rowptr), _p(feature_id), _p(x), _p(target)), "svbfm_set_csr")
        if split == TRAIN:
            self.n_train = n
        elif split == TEST:
            self.n_test = n

    def set_state(self, s):
        wm = np.ascontiguousarray(s["w_mean"], dtype=np.float64)
        wv = np.ascontiguousarray(s["w_var"], dtype=np.float64)
        vm = np.ascontiguousarray(s["v_mean"], dtype=np.float64)
        vv = np.ascontiguousarray(s["v_var"], dtype=np.float64)
        self._ck(lib().svbfm_set_state(self.h, float(s["w0_mean"]), float(s["w0_var"]), _p(wm), _p(wv), _p(vm), _p(vv)),
                 "svbfm_set_state")

    def get_state(self):
        w0m, w0v = C.c_double(), C.c_double()
        wm, wv = np.zeros(self.D), np.zeros(self.D)
        vm, vv = np.zeros((self.K, self.D)), np.zeros((self.K, self.D))
        self._ck(lib().svbfm_get_state(self.h, C.byref(w0m), C.byref(w0v), _p(wm), _p(wv), _p(vm), _p(vv)), "svbfm_get_state")
        return dict(w0_mean=w0m.value, w0_var=w0v.value, w_mean=wm, w_var=wv, v_mean=vm, v_var=vv)

    def get_hyper(self):
        a, s0 = C.c_double(), C.c_double()
        sw, sv = np.zeros(self.G), np.zeros((self.G, max(self.K, 1)))
        self._ck(lib().svbfm_get_hyper(self.h, C.byref(a), C.byref(s0), _p(sw), _p(sv)), "svbfm_get_hyper")
        return dict(alpha=a.value, sigma_0=s0.value, sigma_w=sw, sigma_v=sv[:, :self.K])

    def begin(self):
        self._ck(lib().svbfm_begin(self.h), "svbfm_begin")

    def copies_max_diff(self):
        d = C.c_double(0.0)
        self._ck(lib().svbfm_copies_max_diff(self.h, C.byref(d)), "svbfm_copies_max_diff")
        return d.value

    def reset(self):
        self._ck(lib().svbfm_reset(self.h), "svbfm_reset")

    def run(self, n_iter):
        out = (IterStats * n_iter)()
        self._ck(lib().svbfm_run(self.h, n_iter, out), "svbfm_run")
        return list(out)

    def sweep(self):
        s = IterStats()
        fn = lib().svbfm_vb_sweep if self.method == VB else lib().svbfm_mcmc_sweep
        self._ck(fn(self.h, C.byref(s)), "svbfm_sweep")
        return s

    def vb_online_epoch(self, batch_of_case, num_batch):
        b = np.ascontiguousarray(batch_of_case, dtype=np.uint32)
        s = IterStats()
        self._ck(lib().svbfm_vb_online_epoch(self.h, _p(b), int(num_batch), C.byref(s)), "svbfm_vb_online_epoch")
        return s

    def predict(self, split=TEST):
        out = np.zeros(self.n_test if split == TEST else self.n_train)
        self._ck(lib().svbfm_predict(self.h, split, _p(out)), "svbfm_predict")
        return out

    def get_residuals(self):
        e = np.zeros(self.n_train)
        self._ck(lib().svbfm_get_residuals(self.h, _p(e)), "svbfm_get_residuals")
        return e

    def set_residuals(self, e, sum_t):
        e = np.ascontiguousarray(e, dtype=np.float64)
        self._ck(lib().svbfm_set_residuals(self.h, _p(e), float(sum_t)), "svbfm_set_residuals")

    def set_hyper(self, alpha, sigma_0, sigma_w=None, sigma_v=None):
        sw = np.ascontiguousarray(sigma_w, dtype=np.float64) if sigma_w is not None else None
        sv_ = np.ascontiguousarray(sigma_v, dtype=np.float64) if sigma_v is not None else None
        self._ck(lib().svbfm_set_hyper(self.h, float(alpha), float(sigma_0), _p(sw), _p(sv_)), "svbfm_set_hyper")

    def get_sum_t(self):
        v = C.c_double()
        self._ck(lib().svbfm_get_sum_t(self.h, C.byref(v)), "svbfm_get_sum_t")
        return v.value

    def set_profile(self, on):
        self._ck(lib().svbfm_set_profile(self.h, int(on)), "svbfm_set_profile")

    def get_profile(self):
        names = ["reduce_v", "finalize_v", "apply_v", "reduce_w", "finalize_w", "apply_w", "stream_v_field0", "stream_flush", "stream_v_field1", "stream_w", "collectives", "block_exchange"]
        ms = np.zeros(len(names))     # SVBFM_PROFILE_CLASSES
        cnt = np.zeros(len(names), dtype=np.uint64)
        self._ck(lib().svbfm_get_profile(self.h, _p(ms), _p(cnt)), "svbfm_get_profile")
        return {n: dict(ms=float(m), launches=int(c)) for n, m, c in zip(names, ms, cnt)}

    def info(self):
        i = Info()
        self._ck(lib().svbfm_get_info(self.h, C.byref(i)), "svbfm_get_info")
        return {k: getattr(i, k) for k, _ in Info._fields_}

    def close(self):
        if getattr(self, "h", None):
            lib().svbfm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FmModel:
    """fm_model (reference src/fm_core/fm_model.h:35-64): the fields main() sets (libfm.cpp:259-274)."""

    def __init__(self, num_attribute, num_factor=8, k0=True, k1=True, init_stdev=0.1):
        self.num_attribute, self.num_factor, self.k0, self.k1, self.init_stdev = num_attribute, num_factor, k0, k1, init_stdev
        self.reg0 = self.regw = self.regv = 0.0


class FmLearn:
    """`class fm_learn` (fm_learn.h:38-265): the caller sets fm, meta (attr_group), min/max_target, task, then
    init() and learn(train, test). Per-iteration outputs are returned (and kept in .history) instead of being
    written to files in the CWD; the C++ CLI (host/libfm.cpp) writes the reference's files."""

    method = None

    def __init__(self):
        self.fm = None
        self.attr_group = None          # DataMetaInfo::attr_group
        self.min_target = self.max_target = 0.0
        self.task = 0
        self.num_iter = 100
        self.seed = 42                  # the reference uses time(NULL) (libfm.cpp:123)
        self.device = 0
        self.flags = 0
        self.tile_entries = 0
        self.history = []
        self.engine = None
        self._state = None
        self.comm = None                # (unique_id, rank, world_size)

    def init(self):
        if self.task not in (0, 1):
            raise SvbfmError("unknown task")         # fm_learn.h:87, 111
        if self.task == 1 and self.method != "mcmc":
            raise SvbfmError("classification (task 1) is on the CUDA path for mcmc / als only")
        m = METHODS[self.method]
        self._state = host_init_state(self.seed, self.fm.num_attribute, self.fm.num_factor, self.fm.init_stdev, m)

    def _make_engine(self, **kw):
        fm = self.fm
        self.engine = Engine(self.method, fm.num_attribute, fm.num_factor, fm.k0, fm.k1, self.min_target, self.max_target,
                             device=self.device, seed=self.seed, reg=(fm.reg0, fm.regw, fm.regv), flags=self.flags,
                             tile_entries=self.tile_entries, task=self.task, **kw)
        if self.comm is not None:
            self.engine.comm_init(*self.comm)
        if self.attr_group is not None:
            self.engine.set_groups(self.attr_group)

    def learn(self, train, test):
        self._make_engine()
        E = self.engine
        E.set_csc(TRAIN, train)
        E.set_csc(TEST, test)
        E.set_state(self._state)
        E.begin()
        self.history = E.run(self.num_iter)
        return self.history

    def predict(self, data=None):
        return self.engine.predict(TEST)


class FmLearnVB(FmLearn):
    """fm_learn_vb_simultaneous (fm_learn_vb_simultaneous.h:15-259)."""
    method = "vb"


class FmLearnMCMC(FmLearn):
    """fm_learn_mcmc_simultaneous (fm_learn_mcmc_simultaneous.h:47-305). task = 1: binary classification; the caller maps the
    targets to -1 / +1 first, as the reference's main() does (libfm.cpp:337-343). The iteration statistics then hold accuracies:
    train_stat = "Train=", test_rmse = "Test=" (running mean), rmse_this = acc_mcmc_this; MAP@k needs the reference's
    hard-coded side file (fm_learn.h:124) and is not computed."""
    method = "mcmc"

    def __init__(self):
        super().__init__()
        self.do_sample = True
        self.do_multilevel = True

    def _make_engine(self, **kw):
        super()._make_engine(do_sample=self.do_sample, do_multilevel=self.do_multilevel, **kw)


class FmLearnVBOnline(FmLearn):
    """fm_learn_vb_online_simultaneous (fm_learn_vb_online_simultaneous.h:18-290). The row -> batch rule and the
    libc shuffle stream of the reference are replayed on the host; batches never touch the disk."""
    method = "vb_online"

    def __init__(self):
        super().__init__()
        self.num_batch = 50                         # libfm.cpp:320

    def learn(self, train, test):
        self._make_engine()
        E = self.engine
        E.set_csc(TRAIN, train)
        E.set_csc(TEST, test)
        E.set_state(self._state)
        E.begin()
        n = train.num_cases
        size_except_last = int(np.ceil(n / self.num_batch))       # vbos.h:56-57
        shuffle = np.arange(1, n + 1, dtype=np.uint32)             # vbos.h:58-62
        self.history = []
        for _ in range(self.num_iter):
            lib().svbfm_host_random_shuffle(_p(shuffle), n)        # vbos.h:74
            batch = (np.ceil(shuffle.astype(np.float64) / size_except_last) - 1).astype(np.uint32)   # vbos.h:93
            self.history.append(E.vb_online_epoch(batch, self.num_batch))
        return self.history
