"""CPU, world_size 2 over gloo: the host-side sharding logic of the multi-GPU path. Each rank takes its contiguous
case shard (dist.shard_csc), forms the per-column sufficient statistics of the w sweep on it, the statistics are
all-reduced, and every rank must arrive at exactly the column sums / posterior a single rank computes on all cases."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import svbfm_b200 as sv
    from helpers import ragged, to_csc
    d = sv.submodule("dist")
    tr, _ = ragged(2000, 10, 50, seed=17)
    full = to_csc(tr)
    r = np.random.default_rng(3)
    e, mu = r.normal(size=full.num_cases), r.normal(size=full.num_feature)
    lo, hi = d.shard_bounds(full.num_cases, rank, world)
    mine = d.shard_csc(full, rank, world)
    assert mine.num_cases == hi - lo and int(mine.colptr[-1]) == int(((full.case_id >= lo) & (full.case_id < hi)).sum())
    for j in range(full.num_feature):          # ascending local case ids inside every column
        seg = mine.case_id[int(mine.colptr[j]):int(mine.colptr[j + 1])]
        assert np.all(np.diff(seg.astype(np.int64)) > 0)
    A, B = d.w_column_sums(mine, e[lo:hi], mu)
    t = torch.from_numpy(np.stack([A, B]))
    dist.all_reduce(t)
    uid = d.broadcast_unique_id(lambda: bytes(range(128)), rank)
    A0, B0 = d.w_column_sums(full, e, mu)
    ok = np.allclose(t[0].numpy(), A0, rtol=1e-12, atol=1e-12) and np.allclose(t[1].numpy(), B0, rtol=0, atol=1e-9) and uid == bytes(range(128))
    # identical on every rank after the allreduce -> identical posterior everywhere
    alpha, sw = 1.3, 0.7
    sg = 1.0 / (sw + alpha * t[1].numpy()); m = sg * alpha * t[0].numpy()
    g = [torch.zeros(2, full.num_feature, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(g, torch.from_numpy(np.stack([m, sg])))
    ok = ok and all(torch.equal(g[0], x) for x in g)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_row_sharding_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def _worker_blocks(rank, world, port, q):
    """Sharding by blocks of the first field: the first field's column statistics are complete on the owner (no
    exchange), the second field's still need the allreduce; every case lands on exactly one rank."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import svbfm_b200 as sv
    from helpers import two_field, to_csc
    d = sv.submodule("dist")
    U = 120
    tr, _ = two_field(5000, 10, U, 90, seed=19)
    full = to_csc(tr)
    r = np.random.default_rng(5)
    e, mu = r.normal(size=full.num_cases), r.normal(size=full.num_feature)
    b = d.block_bounds(full.colptr, U, world)
    mine, ids = d.shard_csc_by_block(full, rank, world, U)
    ok = b[0] == 0 and b[-1] == U and all(x <= y for x, y in zip(b, b[1:]))
    cnt = torch.tensor([mine.num_cases])
    dist.all_reduce(cnt)
    ok = ok and int(cnt) == full.num_cases
    own = np.zeros(full.num_cases)
    own[ids] = 1
    t = torch.from_numpy(own)
    dist.all_reduce(t)
    ok = ok and bool((t == 1).all())                       # every case on exactly one rank
    nz = np.nonzero(np.diff(mine.colptr[:U + 1].astype(np.int64)))[0]
    ok = ok and (len(nz) == 0 or (nz.min() >= b[rank] and nz.max() < b[rank + 1]))   # only columns of the own block
    A, B = d.w_column_sums(mine, e[ids], mu)
    A0, B0 = d.w_column_sums(full, e, mu)
    blk = slice(b[rank], b[rank + 1])
    ok = ok and np.allclose(A[blk], A0[blk], rtol=1e-12, atol=1e-12) and np.array_equal(B[blk], B0[blk])   # complete without exchange
    t = torch.from_numpy(np.stack([A[U:], B[U:]]))
    dist.all_reduce(t)
    ok = ok and np.allclose(t[0].numpy(), A0[U:], rtol=1e-12, atol=1e-12) and np.allclose(t[1].numpy(), B0[U:], rtol=0, atol=1e-9)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_block_sharding_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29100 + os.getpid() % 300
    ps = [ctx.Process(target=_worker_blocks, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
