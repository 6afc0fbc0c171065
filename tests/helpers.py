"""Shared test helpers: seeded synthetic data in both the oracle's (CSR) and the product's (CSC) form."""
import numpy as np

import oracle_binding as ob
import svbfm_b200 as sv

synth = sv.submodule("synth")


def two_field(N, Nt, U, I, seed=1, values=False):
    model = synth.planted_model(U, I, seed + 100)
    u, i, y = synth.ratings(N, U, I, model, seed)
    ut, it, yt = synth.ratings(Nt, U, I, model, seed + 1)
    tr = ob.Csr(*synth.to_csr(u, i, y, U))
    te = ob.Csr(*synth.to_csr(ut, it, yt, U))
    if values:   # non one-hot values exercise the x != 1 kernels
        r = np.random.default_rng(seed + 7)
        tr.val[:] = r.uniform(0.5, 1.5, len(tr.val)).astype(np.float32)
        te.val[:] = r.uniform(0.5, 1.5, len(te.val)).astype(np.float32)
    return tr, te


def ragged(N, Nt, D, seed=3, max_nnz=5, fields=None):
    """General sparse cases: 0..max_nnz distinct features per case with real values (multi-hot, ragged, some empty)."""
    r = np.random.default_rng(seed)

    def make(n, s):
        rr = np.random.default_rng(s)
        lens = rr.integers(0, max_nnz + 1, n)
        rowptr = np.zeros(n + 1, dtype=np.uint64)
        rowptr[1:] = np.cumsum(lens)
        col = np.zeros(int(rowptr[-1]), dtype=np.uint32)
        for k in range(n):
            col[int(rowptr[k]):int(rowptr[k + 1])] = np.sort(rr.choice(D, size=lens[k], replace=False))
        val = rr.uniform(0.2, 2.0, len(col)).astype(np.float32)
        y = rr.integers(1, 6, n).astype(np.float32)
        return ob.Csr(rowptr, col, val, y)
    return make(N, seed), make(Nt, seed + 1)


def to_csc(csr):
    return sv.CscData.from_csr(csr.rowptr, csr.col, csr.val, csr.y, csr.n_feat)


def make_learner(method, tr, te, K, seed=42, k0=1, k1=1, groups=None, num_iter=5, flags=0, tile_entries=0, D=None, **kw):
    cls = {"vb": sv.FmLearnVB, "mcmc": sv.FmLearnMCMC, "vb_online": sv.FmLearnVBOnline}[method]
    L = cls()
    if D is None:
        D = max(tr.n_feat, te.n_feat) + (0 if method == "vb_online" else 1)
    L.fm = sv.FmModel(D, K, bool(k0), bool(k1))
    L.min_target, L.max_target = float(tr.y.min()), float(tr.y.max())
    L.num_iter = num_iter
    L.seed = seed
    L.flags = flags
    L.tile_entries = tile_entries
    L.attr_group = groups
    for k, v in kw.items():
        setattr(L, k, v)
    L.init()
    return L


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)
