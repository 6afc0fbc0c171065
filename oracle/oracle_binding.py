"""ctypes binding of the CPU oracle (oracle/svbfm_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.
The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsvbfm_oracle.so")

VB, VB_ONLINE, MCMC = 0, 1, 2
METHODS = {"vb": VB, "vb_online": VB_ONLINE, "mcmc": MCMC}


class Stats(C.Structure):
    _fields_ = [("test_rmse", C.c_double), ("train_stat", C.c_double), ("free_energy", C.c_double),
                ("alpha", C.c_double), ("rmse_this", C.c_double), ("has_free_energy", C.c_int),
                ("nan_inf_count", C.c_uint32)]


class CsrStruct(C.Structure):
    _fields_ = [("n_rows", C.c_uint32), ("n_feat", C.c_uint32), ("nnz", C.c_uint64),
                ("rowptr", C.POINTER(C.c_uint64)), ("col", C.POINTER(C.c_uint32)),
                ("val", C.POINTER(C.c_float)), ("y", C.POINTER(C.c_float)),
                ("min_target", C.c_float), ("max_target", C.c_float)]


def build(force=False):
    src = os.path.join(_HERE, "svbfm_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "port"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_split.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_set_groups.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_init.argtypes = [C.c_void_p, C.c_long, C.c_double]
        L.orc_set_mcmc_options.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_set_num_batch.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_set_task.argtypes = [C.c_void_p, C.c_int32]
        L.orc_cdf_gaussian.argtypes = [C.c_double]
        L.orc_cdf_gaussian.restype = C.c_double
        L.orc_ran_left_tgaussian.argtypes = [C.c_double] * 3
        L.orc_ran_left_tgaussian.restype = C.c_double
        L.orc_ran_right_tgaussian.argtypes = [C.c_double] * 3
        L.orc_ran_right_tgaussian.restype = C.c_double
        L.orc_set_regular.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        L.orc_begin.argtypes = [C.c_void_p]
        L.orc_iterate.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.orc_get_state.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)] + [C.c_void_p] * 4
        L.orc_set_state.argtypes = [C.c_void_p, C.c_double, C.c_double] + [C.c_void_p] * 4
        L.orc_get_hyper.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p, C.c_void_p]
        L.orc_get_train_cache.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_get_test_pred.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_parse_text.argtypes = [C.c_char_p, C.POINTER(CsrStruct)]
        L.orc_csr_free.argtypes = [C.POINTER(CsrStruct)]
        L.orc_transpose.argtypes = [C.POINTER(CsrStruct), C.c_uint32, C.POINTER(CsrStruct)]
        L.orc_write_x.argtypes = [C.c_char_p, C.POINTER(CsrStruct), C.c_uint32]
        L.orc_write_y.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32]
        L.orc_read_x.argtypes = [C.c_char_p, C.POINTER(CsrStruct)]
        for f in ("orc_ran_uniform", "orc_ran_gaussian"):
            getattr(L, f).restype = C.c_double
        L.orc_ran_gamma.restype = C.c_double
        L.orc_ran_gamma.argtypes = [C.c_double]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Csr:
    """Host CSR of a libFM data set (rows = cases)."""

    def __init__(self, rowptr, col, val, y, n_feat=None):
        self.rowptr = np.ascontiguousarray(rowptr, dtype=np.uint64)
        self.col = np.ascontiguousarray(col, dtype=np.uint32)
        self.val = np.ascontiguousarray(val, dtype=np.float32)
        self.y = np.ascontiguousarray(y, dtype=np.float32)
        self.n_rows = len(self.y)
        self.n_feat = int(n_feat) if n_feat is not None else (int(self.col.max()) + 1 if len(self.col) else 0)

    def as_struct(self):
        s = CsrStruct()
        s.n_rows, s.n_feat, s.nnz = self.n_rows, self.n_feat, len(self.col)
        s.rowptr = self.rowptr.ctypes.data_as(C.POINTER(C.c_uint64))
        s.col = self.col.ctypes.data_as(C.POINTER(C.c_uint32))
        s.val = self.val.ctypes.data_as(C.POINTER(C.c_float))
        s.y = self.y.ctypes.data_as(C.POINTER(C.c_float))
        return s


def _from_struct(s, has_y=True):
    n, nnz = s.n_rows, s.nnz
    rowptr = np.ctypeslib.as_array(s.rowptr, shape=(n + 1,)).copy()
    col = np.ctypeslib.as_array(s.col, shape=(max(nnz, 1),))[:nnz].copy()
    val = np.ctypeslib.as_array(s.val, shape=(max(nnz, 1),))[:nnz].copy()
    y = np.ctypeslib.as_array(s.y, shape=(max(n, 1),))[:n].copy() if has_y and s.y else np.zeros(n, np.float32)
    return Csr(rowptr, col, val, y, n_feat=s.n_feat)


def parse_text(path):
    s = CsrStruct()
    r = lib().orc_parse_text(path.encode(), C.byref(s))
    if r:
        raise RuntimeError(f"oracle parse_text({path}) failed: {r}")
    out = _from_struct(s)
    lib().orc_csr_free(C.byref(s))
    return out


def transpose(csr, n_out_rows):
    """Counting transpose; returns (ptr, id, val) of the CSC (rows = features)."""
    s = csr.as_struct()
    o = CsrStruct()
    r = lib().orc_transpose(C.byref(s), n_out_rows, C.byref(o))
    if r:
        raise RuntimeError(f"oracle transpose failed: {r}")
    t = _from_struct(o, has_y=False)
    lib().orc_csr_free(C.byref(o))
    return t.rowptr, t.col, t.val


def write_x(path, csr, num_cols):
    s = csr.as_struct()
    assert lib().orc_write_x(path.encode(), C.byref(s), num_cols) == 0


def write_y(path, y):
    y = np.ascontiguousarray(y, dtype=np.float32)
    assert lib().orc_write_y(path.encode(), _p(y), len(y)) == 0


def read_x(path):
    s = CsrStruct()
    r = lib().orc_read_x(path.encode(), C.byref(s))
    if r:
        raise RuntimeError(f"oracle read_x({path}) failed: {r}")
    out = _from_struct(s, has_y=False)
    lib().orc_csr_free(C.byref(s))
    return out


class Oracle:
    """One learner instance of the CPU restatement (vb | vb_online | mcmc)."""

    def __init__(self, method, train, test, K, k0=1, k1=1, D=None, seed=42, init_stdev=0.1,
                 groups=None, num_batch=None, do_sample=True, do_multilevel=True, reg=None, task=0):
        L = lib()
        self.method = METHODS[method] if isinstance(method, str) else method
        if D is None:
            if self.method == VB_ONLINE:   # libfm.cpp:528-599 keeps the max id, then +1 at :215
                D = max(train.n_feat, test.n_feat)
            else:                          # libfm.cpp:215 (fork-specific "+1")
                D = max(train.n_feat, test.n_feat) + 1
        self.D, self.K = int(D), int(K)
        self.train, self.test = train, test
        self.h = L.orc_create(self.method, self.D, self.K, int(k0), int(k1))
        tr_feat = self.D if self.method == VB_ONLINE else train.n_feat
        assert L.orc_set_split(self.h, 0, train.n_rows, tr_feat, _p(train.rowptr), _p(train.col), _p(train.val), _p(train.y)) == 0
        assert L.orc_set_split(self.h, 1, test.n_rows, test.n_feat, _p(test.rowptr), _p(test.col), _p(test.val), _p(test.y)) == 0
        self.G = 1
        if groups is not None:
            g = np.ascontiguousarray(groups, dtype=np.uint32)
            assert len(g) == self.D
            self.G = int(g.max()) + 1
            assert L.orc_set_groups(self.h, _p(g), self.G) == 0
        if num_batch is not None:
            L.orc_set_num_batch(self.h, int(num_batch))
        L.orc_set_mcmc_options(self.h, int(do_sample), int(do_multilevel))
        if task:
            assert L.orc_set_task(self.h, int(task)) == 0
        assert L.orc_init(self.h, int(seed), float(init_stdev)) == 0
        if reg is not None:
            assert L.orc_set_regular(self.h, float(reg[0]), float(reg[1]), float(reg[2])) == 0
        self._begun = False

    def begin(self):
        assert lib().orc_begin(self.h) == 0
        self._begun = True

    def iterate(self):
        if not self._begun:
            self.begin()
        s = Stats()
        assert lib().orc_iterate(self.h, C.byref(s)) == 0
        return s

    def get_state(self):
        w0m, w0v = C.c_double(), C.c_double()
        wm, wv = np.zeros(self.D), np.zeros(self.D)
        vm, vv = np.zeros((self.K, self.D)), np.zeros((self.K, self.D))
        lib().orc_get_state(self.h, C.byref(w0m), C.byref(w0v), _p(wm), _p(wv), _p(vm), _p(vv))
        return dict(w0_mean=w0m.value, w0_var=w0v.value, w_mean=wm, w_var=wv, v_mean=vm, v_var=vv)

    def get_hyper(self):
        a, s0 = C.c_double(), C.c_double()
        sw, sv = np.zeros(self.G), np.zeros((self.G, self.K))
        lib().orc_get_hyper(self.h, C.byref(a), C.byref(s0), _p(sw), _p(sv))
        return dict(alpha=a.value, sigma_0=s0.value, sigma_w=sw, sigma_v=sv)

    def get_train_cache(self, want_t=True):
        e = np.zeros(self.train.n_rows)
        t = np.zeros(self.train.n_rows) if want_t else None
        lib().orc_get_train_cache(self.h, _p(e), _p(t))
        return e, t

    def get_test_pred(self):
        p = np.zeros(self.test.n_rows)
        lib().orc_get_test_pred(self.h, _p(p))
        return p

    def close(self):
        if self.h:
            lib().orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
