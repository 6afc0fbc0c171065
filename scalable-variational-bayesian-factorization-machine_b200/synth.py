"""Synthetic MovieLens-shaped ratings: every case has exactly two non-zeros `u:1 (U+i):1`,
users [0,U), items [U,U+I), Zipf(1.0) popularity over a random permutation, cases in random order,
y = clip(round(3.5 + b_u + b_i + <p_u,q_i> + eps), 1, 5) with a planted rank-8 model."""
import numpy as np

SHAPES = {   # name: (U, I, N, Nt, K)
    "ml1m": (6040, 3952, 1_000_000, 100_000, 20),
    "ml10m": (71_567, 10_681, 10_000_000, 1_000_000, 50),
    "netflix": (480_189, 17_770, 100_000_000, 1_400_000, 100),
    "kdd200m": (1_000_990, 624_961, 200_000_000, 2_000_000, 50),
}


def _zipf_p(n):
    p = 1.0 / np.arange(1, n + 1, dtype=np.float64)
    return p / p.sum()


def planted_model(U, I, seed, rank=8):
    r = np.random.default_rng(seed)
    return dict(bu=r.normal(0, 0.4, U), bi=r.normal(0, 0.4, I), P=r.normal(0, 0.3, (U, rank)), Q=r.normal(0, 0.3, (I, rank)),
                perm_u=r.permutation(U), perm_i=r.permutation(I))


def ratings(n, U, I, model, seed):
    """numpy generator (host). Returns (user[n], item[n], y[n] float32)."""
    r = np.random.default_rng(seed)
    u = model["perm_u"][r.choice(U, size=n, p=_zipf_p(U))]
    i = model["perm_i"][r.choice(I, size=n, p=_zipf_p(I))]
    s = 3.5 + model["bu"][u] + model["bi"][i] + np.einsum("nk,nk->n", model["P"][u], model["Q"][i]) + r.normal(0, 0.8, n)
    y = np.clip(np.round(s), 1, 5).astype(np.float32)
    return u.astype(np.uint32), i.astype(np.uint32), y


def to_csr(u, i, y, U):
    n = len(y)
    rowptr = np.arange(0, 2 * n + 1, 2, dtype=np.uint64)
    col = np.empty(2 * n, dtype=np.uint32)
    col[0::2] = u
    col[1::2] = U + i
    return rowptr, col, np.ones(2 * n, dtype=np.float32), y


def write_libfm_text(path, u, i, y, U):
    with open(path, "w") as f:
        for a, b, c in zip(y, u, i):
            f.write(f"{int(a)} {int(b)}:1 {int(U + c)}:1\n")


def csc_two_field(u, i, y, U, I, num_cols=None):
    """CSC (colptr, case_id, x) of the two one-hot fields, case ids ascending inside every column."""
    n = len(y)
    nc = int(num_cols) if num_cols is not None else U + I
    ou = np.argsort(u, kind="stable")
    oi = np.argsort(i, kind="stable")
    colptr = np.zeros(nc + 1, dtype=np.uint64)
    cnt = np.zeros(nc, dtype=np.int64)
    cnt[:U] = np.bincount(u, minlength=U)
    cnt[U:U + I] = np.bincount(i, minlength=I)
    np.cumsum(cnt, out=colptr[1:])
    case_id = np.concatenate([ou, oi]).astype(np.uint32)
    return colptr, case_id, np.ones(2 * n, dtype=np.float32)
