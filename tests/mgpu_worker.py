"""Worker of tests/test_gpu_multi.py: launched by torch.distributed.run with one process per GPU. Every rank
feeds its contiguous shard of the cases through the C-ABI; rank 0 compares the per-iteration statistics with the
CPU oracle run on ALL cases (tolerance 1e-7: only the summation order differs) and prints MGPU_OK."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist

import oracle_binding as ob
import svbfm_b200 as sv
from helpers import ragged, rel, to_csc, two_field


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d = sv.submodule("dist")

    def get_id():
        buf = (sv.C.c_uint8 * sv.COMM_ID_BYTES)()
        assert sv.lib().svbfm_comm_get_unique_id(buf) == 0
        return bytes(buf)

    ok = True
    for name, (tr, te), method, K, iters in (("two_field_vb", two_field(30000, 3000, 400, 300, seed=51), "vb", 4, 5),
                                             ("two_field_vb_blocks", two_field(30000, 3000, 400, 300, seed=54), "vb", 4, 5),
                                             ("ragged_vb", ragged(4000, 500, 60, seed=52), "vb", 3, 4),
                                             ("two_field_als", two_field(20000, 2000, 300, 200, seed=53), "mcmc", 3, 4),
                                             ("two_field_als_blocks", two_field(20000, 2000, 300, 200, seed=55), "mcmc", 3, 4),
                                             # cross shards: first residual copy by user block, second copy by item block (records allgathered)
                                             ("two_field_vb_cross", two_field(30000, 3000, 400, 300, seed=58), "vb", 4, 5),
                                             ("two_field_als_cross", two_field(20000, 2000, 300, 200, seed=59), "mcmc", 3, 4),
                                             ("two_field_vbo", two_field(20000, 2000, 300, 200, seed=56), "vb_online", 3, 3)):
        uid = d.broadcast_unique_id(get_id, rank, device=torch.device("cuda", local))
        D = max(tr.n_feat, te.n_feat) + (0 if method == "vb_online" else 1)
        kw = dict(do_sample=False, do_multilevel=False) if method == "mcmc" else {}
        E = sv.Engine(method, D, K, 1, 1, float(tr.y.min()), float(tr.y.max()), device=local, seed=42, **kw)
        E.comm_init(uid, rank, world)
        cross = name.endswith("_cross")
        blocks = name.endswith("_blocks") or cross       # cases sharded by blocks of the first field (users): no exchange for that field
        nu = 400 if "vb" in name else 300
        E.set_csc(sv.TRAIN, d.shard_csc_by_block(to_csc(tr), rank, world, nu)[0] if blocks else d.shard_csc(to_csc(tr), rank, world))
        if cross:
            E.set_csc(sv.TRAIN_SECOND, d.shard_csc_by_second_block(to_csc(tr), rank, world, nu)[0])
        E.set_csc(sv.TEST, d.shard_csc(to_csc(te), rank, world))
        assert E.info()["exclusive_blocks"] & 3 == (3 if cross else 1 if blocks else 0), (name, E.info())
        if name.startswith("two_field"):
            assert E.info()["fused_schedule"] & 1
        E.set_state(sv.host_init_state(42, D, K, 0.1, sv.METHODS[method]))
        E.begin()
        okw = {}
        if method == "vb_online":      # sharded vb_online on the stream schedule: the case -> batch rule is replayed on the libc stream (vbos.h:74-95)
            nb, n = 5, tr.n_rows
            okw = dict(num_batch=nb)
            size_except_last = int(np.ceil(n / nb))
            shuffle = np.arange(1, n + 1, dtype=np.uint32)
            lo, hi = d.shard_bounds(n, rank, world)
            hist = []
            for _ in range(iters):
                sv.lib().svbfm_host_random_shuffle(shuffle.ctypes.data_as(sv.C.c_void_p), n)
                batch = (np.ceil(shuffle.astype(np.float64) / size_except_last) - 1).astype(np.uint32)
                hist.append(E.vb_online_epoch(np.ascontiguousarray(batch[lo:hi]), nb))
        else:
            hist = E.run(iters)
        if rank == 0:
            orc = ob.Oracle(method, tr, te, K=K, seed=42, **kw, **okw)
            for it, s in enumerate(hist):
                o = orc.iterate()
                good = rel(s.test_rmse, o.test_rmse) < 1e-7 and (method == "vb_online" or rel(s.train_stat, o.train_stat) < 1e-7)
                if method != "mcmc":
                    good = good and rel(s.free_energy, o.free_energy) < 1e-7
                if not good:
                    print("MISMATCH", name, it, s.test_rmse, o.test_rmse, s.free_energy, o.free_energy, flush=True)
                ok = ok and good
        # replicated parameters are bit-identical on every rank
        st = E.get_state()
        t = torch.from_numpy(np.concatenate([st["w_mean"], st["v_mean"].ravel()])).cuda()
        mx, mn = t.clone(), t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        ok = ok and bool(torch.equal(mx, mn))
        assert E.info()["world_size"] == world
        if method != "vb_online":
            ok = ok and E.copies_max_diff() == 0.0
        E.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MGPU_OK" if int(flag) == 1 else "MGPU_FAIL", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
