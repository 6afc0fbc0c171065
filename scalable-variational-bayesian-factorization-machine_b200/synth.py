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


# ---- torch generator for the large configs (runs on the GPU; same distributions as above) -----------------
def _model_torch(U, I, device, model_seed):
    import torch
    gm = torch.Generator(device=device)
    gm.manual_seed(model_seed)
    bu = torch.randn(U, generator=gm, device=device) * 0.4
    bi = torch.randn(I, generator=gm, device=device) * 0.4
    P = torch.randn(U, 8, generator=gm, device=device) * 0.3
    Q = torch.randn(I, 8, generator=gm, device=device) * 0.3
    perm_u = torch.randperm(U, generator=gm, device=device)
    perm_i = torch.randperm(I, generator=gm, device=device)
    return bu, bi, P, Q, perm_u, perm_i


def user_mass_torch(U, I, device, model_seed=20261017):
    """Expected share of the ratings of every user id under ratings_torch (Zipf(1) over the planted permutation)."""
    import torch
    perm_u = _model_torch(U, I, device, model_seed)[4]
    p = 1.0 / torch.arange(1, U + 1, device=device, dtype=torch.float64)
    mass = torch.empty(U, device=device, dtype=torch.float64)
    mass[perm_u] = p / p.sum()
    return mass


def ratings_torch(n, U, I, seed, device, chunk=25_000_000, model_seed=20261017, user_range=None):
    """Returns (user int32[m], item int32[m], y float32[m]) on `device`; deterministic for a given seed. m = n, or, with
    user_range = (lo, hi), the subset of the same n draws whose user id lies in [lo, hi) (a user-block shard)."""
    import torch
    bu, bi, P, Q, perm_u, perm_i = _model_torch(U, I, device, model_seed)
    cdf_u = torch.cumsum(1.0 / torch.arange(1, U + 1, device=device, dtype=torch.float64), 0)
    cdf_u /= cdf_u[-1].clone()
    cdf_i = torch.cumsum(1.0 / torch.arange(1, I + 1, device=device, dtype=torch.float64), 0)
    cdf_i /= cdf_i[-1].clone()
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    us, its, ys = [], [], []
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        ru = torch.searchsorted(cdf_u, torch.rand(b - a, generator=g, device=device, dtype=torch.float64)).clamp_(max=U - 1)
        ri = torch.searchsorted(cdf_i, torch.rand(b - a, generator=g, device=device, dtype=torch.float64)).clamp_(max=I - 1)
        uu, ii = perm_u[ru], perm_i[ri]
        s = 3.5 + bu[uu] + bi[ii] + (P[uu] * Q[ii]).sum(1) + 0.8 * torch.randn(b - a, generator=g, device=device)
        yy = s.round_().clamp_(1, 5)
        if user_range is not None:
            keep = (uu >= user_range[0]) & (uu < user_range[1])
            uu, ii, yy = uu[keep], ii[keep], yy[keep]
        us.append(uu.to(torch.int32)); its.append(ii.to(torch.int32)); ys.append(yy.to(torch.float32))
    return torch.cat(us), torch.cat(its), torch.cat(ys)


def csc_two_field_torch(u, it, U, I, num_cols=None):
    """CSC of the two one-hot fields built with torch sorts on the device. Returns (colptr int64[nc+1], case_id int32[2n])."""
    import torch
    nc = int(num_cols) if num_cols is not None else U + I
    n = u.numel()
    case_id = torch.empty(2 * n, dtype=torch.int32, device=u.device)
    case_id[:n] = torch.sort(u, stable=True)[1].to(torch.int32)
    case_id[n:] = torch.sort(it, stable=True)[1].to(torch.int32)
    cnt = torch.zeros(nc, dtype=torch.int64, device=u.device)
    cnt[:U] = torch.bincount(u, minlength=U)
    cnt[U:U + I] = torch.bincount(it, minlength=I)
    colptr = torch.zeros(nc + 1, dtype=torch.int64, device=u.device)
    colptr[1:] = torch.cumsum(cnt, 0)
    return colptr, case_id
