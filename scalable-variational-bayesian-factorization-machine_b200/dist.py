"""Host-side logic of the row-sharded multi-GPU path: one process per GPU, cases split
into contiguous shards, features and parameters replicated. The only exchange on the data path is the
allreduce of the per-run column sums, done inside the engine by NCCL; this module holds what the host does:
shard bounds, shard extraction from a CSC data set and the distribution of the NCCL unique id."""
import numpy as np


def shard_bounds(n, rank, world):
    """Contiguous, balanced case ranges: [n*rank/world, n*(rank+1)/world)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def shard_csc(data, rank, world):
    """Cases [lo, hi) of a CscData with local case ids (ascending order inside every column is preserved)."""
    import sys
    CscData = sys.modules[__name__.rsplit('.', 1)[0]].CscData   # the package is loaded by path (svbfm_b200.py)
    lo, hi = shard_bounds(data.num_cases, rank, world)
    keep = (data.case_id >= lo) & (data.case_id < hi)
    col_of = np.repeat(np.arange(data.num_feature, dtype=np.int64), np.diff(data.colptr.astype(np.int64)))
    colptr = np.zeros(data.num_feature + 1, dtype=np.uint64)
    np.cumsum(np.bincount(col_of[keep], minlength=data.num_feature), out=colptr[1:])
    return CscData(colptr, (data.case_id[keep] - lo).astype(np.uint32), (data.x[keep] if data.x is not None else None), data.target[lo:hi])


def block_bounds(colptr, first_field_cols, world):
    """Column boundaries [world + 1] that split the first field's columns into rank-ordered blocks with about the same
    number of entries each ('partition rows by user block')."""
    cum = np.asarray(colptr[:first_field_cols + 1], dtype=np.int64)
    total = int(cum[-1])
    b = [0]
    for r in range(1, world):
        b.append(int(np.searchsorted(cum, total * r // world, side="left")))
    b.append(first_field_cols)
    return [min(max(x, 0), first_field_cols) for x in b]


def shard_csc_by_block(data, rank, world, first_field_cols):
    """Cases whose first-field feature lies in this rank's column block (block_bounds), local case ids in ascending
    original order. The engine detects the disjoint blocks and updates the first field without any exchange.
    Returns (CscData, original case ids of the shard)."""
    import sys
    CscData = sys.modules[__name__.rsplit('.', 1)[0]].CscData
    b = block_bounds(data.colptr, first_field_cols, world)
    cp = data.colptr.astype(np.int64)
    mine = np.sort(data.case_id[cp[b[rank]]:cp[b[rank + 1]]].astype(np.int64))     # one first-field entry per case
    local = np.full(data.num_cases, -1, dtype=np.int64)
    local[mine] = np.arange(len(mine))
    keep = local[data.case_id] >= 0
    col_of = np.repeat(np.arange(data.num_feature, dtype=np.int64), np.diff(cp))
    colptr = np.zeros(data.num_feature + 1, dtype=np.uint64)
    np.cumsum(np.bincount(col_of[keep], minlength=data.num_feature), out=colptr[1:])
    return CscData(colptr, local[data.case_id[keep]].astype(np.uint32), (data.x[keep] if data.x is not None else None), data.target[mine]), mine


def second_block_bounds(colptr, first_field_cols, num_cols, world):
    """Column boundaries [world + 1] that split the SECOND field's columns [first_field_cols, num_cols) into rank-ordered blocks
    with about the same number of entries each (cross shards: the second residual copy is sharded by these blocks)."""
    cum = np.asarray(colptr[first_field_cols:num_cols + 1], dtype=np.int64) - int(colptr[first_field_cols])
    total = int(cum[-1])
    b = [0]
    for r in range(1, world):
        b.append(int(np.searchsorted(cum, total * r // world, side="left")))
    b.append(num_cols - first_field_cols)
    return [first_field_cols + min(max(x, 0), num_cols - first_field_cols) for x in b]


def shard_csc_by_second_block(data, rank, world, first_field_cols):
    """Cases whose second-field feature lies in this rank's block of the second field's columns: the shard a rank hands over as
    split TRAIN_SECOND after its first-field block (shard_csc_by_block) went in as TRAIN. Returns (CscData, original case ids)."""
    import sys
    CscData = sys.modules[__name__.rsplit('.', 1)[0]].CscData
    b = second_block_bounds(data.colptr, first_field_cols, data.num_feature, world)
    cp = data.colptr.astype(np.int64)
    mine = np.sort(data.case_id[cp[b[rank]]:cp[b[rank + 1]]].astype(np.int64))     # one second-field entry per case
    local = np.full(data.num_cases, -1, dtype=np.int64)
    local[mine] = np.arange(len(mine))
    keep = local[data.case_id] >= 0
    col_of = np.repeat(np.arange(data.num_feature, dtype=np.int64), np.diff(cp))
    colptr = np.zeros(data.num_feature + 1, dtype=np.uint64)
    np.cumsum(np.bincount(col_of[keep], minlength=data.num_feature), out=colptr[1:])
    return CscData(colptr, local[data.case_id[keep]].astype(np.uint32), (data.x[keep] if data.x is not None else None), data.target[mine]), mine


def broadcast_unique_id(get_id, rank, device=None):
    """Rank 0 calls get_id() -> bytes; every rank returns the same bytes (torch.distributed broadcast; gloo or nccl)."""
    import torch
    import torch.distributed as dist
    buf = torch.zeros(128, dtype=torch.uint8, device=device if device is not None else "cpu")
    if rank == 0:
        buf.copy_(torch.tensor(list(get_id()), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().tolist())


def w_column_sums(data, e, mu_w):
    """Per-column sufficient statistics of the w sweep on a shard: A_j = sum x (e_i + x mu_j), B_j = sum x^2
    (fm_learn_vb.h:534-539). numpy, used by the gloo test of the sharding algebra."""
    col_of = np.repeat(np.arange(data.num_feature, dtype=np.int64), np.diff(data.colptr.astype(np.int64)))
    x = data.x.astype(np.float64)
    A = np.bincount(col_of, weights=x * (e[data.case_id] + x * mu_w[col_of]), minlength=data.num_feature)
    B = np.bincount(col_of, weights=x * x, minlength=data.num_feature)
    return A, B
