"""TEST INFRASTRUCTURE: one rank of the emulated multi-rank run (tests/test_emu_kernels.py). The same cases as
tests/mgpu_worker.py (the 2-GPU test), but the engine is the host build of the kernels (libsvbfm_emu.so) and the collectives
go through tests/emu/fake_nccl.c; no torch. usage: mrank_worker.py <rank> <world> <scratch dir> [case filter]"""
import hashlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]

import numpy as np

import oracle_binding as ob
import svbfm_b200 as sv
from helpers import ragged, rel, to_csc, two_field


def exchange(scratch, tag, rank, world, payload=None):
    """rank 0 publishes `payload` (bytes) under `tag`; every rank returns it"""
    path = os.path.join(scratch, tag)
    if rank == 0:
        with open(path + ".tmp", "wb") as f:
            f.write(payload)
        os.rename(path + ".tmp", path)
        return payload
    for _ in range(60000):
        if os.path.exists(path):
            return open(path, "rb").read()
        time.sleep(0.005)
    raise RuntimeError("rank 0 never published " + tag)


def main():
    rank, world, scratch = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    only = sys.argv[4] if len(sys.argv) > 4 else ""
    d = sv.submodule("dist")
    ok = True
    cases = (("two_field_vb", two_field(12000, 1200, 200, 150, seed=51), "vb", 3, 3, {}),
             ("two_field_vb_blocks", two_field(12000, 1200, 200, 150, seed=54), "vb", 3, 3, {}),
             ("ragged_vb", ragged(3000, 400, 60, seed=52), "vb", 2, 3, {}),
             ("two_field_als", two_field(10000, 1000, 160, 120, seed=53), "mcmc", 2, 3, {}),
             ("two_field_als_blocks", two_field(10000, 1000, 160, 120, seed=55), "mcmc", 2, 3, {}),
             # cross shards: first copy by user block, second copy by item block (svbfm_set_csc(TRAIN_SECOND))
             ("two_field_vb_cross", two_field(12000, 1200, 200, 150, seed=58), "vb", 3, 3, {}),
             ("two_field_als_cross", two_field(10000, 1000, 160, 120, seed=59), "mcmc", 2, 3, {}),
             ("two_field_vbo", two_field(10000, 1000, 160, 120, seed=56), "vb_online", 2, 2, dict(num_batch=4)),
             # three cases per batch: some ranks have no case of a batch and still have to take part in its collectives
             ("two_field_vbo_tiny_batches", two_field(300, 100, 20, 15, seed=57), "vb_online", 2, 2, dict(num_batch=100)))
    for name, (tr, te), method, K, iters, extra in cases:
        if only and only not in name:
            continue
        buf = (sv.C.c_uint8 * sv.COMM_ID_BYTES)()
        if rank == 0:
            assert sv.lib().svbfm_comm_get_unique_id(buf) == 0
        uid = exchange(scratch, f"uid_{name}", rank, world, bytes(buf))
        D = max(tr.n_feat, te.n_feat) + (0 if method == "vb_online" else 1)
        kw = dict(do_sample=False, do_multilevel=False) if method == "mcmc" else {}
        E = sv.Engine(method, D, K, 1, 1, float(tr.y.min()), float(tr.y.max()), seed=42, tile_entries=64, **kw)
        E.comm_init(uid, rank, world)
        cross = name.endswith("_cross")
        blocks = name.endswith("_blocks") or cross
        nu = 200 if name.startswith("two_field_vb") else 160
        if blocks:
            shard, mine = d.shard_csc_by_block(to_csc(tr), rank, world, nu)
        else:
            shard = d.shard_csc(to_csc(tr), rank, world)
            lo, hi = d.shard_bounds(tr.n_rows, rank, world)
            mine = np.arange(lo, hi)
        E.set_csc(sv.TRAIN, shard)
        if cross:
            E.set_csc(sv.TRAIN_SECOND, d.shard_csc_by_second_block(to_csc(tr), rank, world, nu)[0])
        E.set_csc(sv.TEST, d.shard_csc(to_csc(te), rank, world))
        info = E.info()
        assert info["exclusive_blocks"] & 3 == (3 if cross else 1 if blocks else 0), (name, info)
        assert info["world_size"] == world
        E.set_state(sv.host_init_state(42, D, K, 0.1, sv.METHODS[method]))
        E.begin()
        if method == "vb_online":
            nb = extra["num_batch"]
            n = tr.n_rows
            size_except_last = int(np.ceil(n / nb))
            shuffle = np.arange(1, n + 1, dtype=np.uint32)
            sv.lib().svbfm_host_init_state      # (the libc stream continues after host_init_state, like in the reference)
            hist = []
            for _ in range(iters):
                sv.lib().svbfm_host_random_shuffle(shuffle.ctypes.data_as(sv.C.c_void_p), n)
                batch = (np.ceil(shuffle.astype(np.float64) / size_except_last) - 1).astype(np.uint32)
                hist.append(E.vb_online_epoch(np.ascontiguousarray(batch[mine]), nb))
        else:
            hist = E.run(iters)
        if rank == 0:
            orc = ob.Oracle(method, tr, te, K=K, seed=42, **kw, **extra)
            for it, s in enumerate(hist):
                o = orc.iterate()
                good = rel(s.test_rmse, o.test_rmse) < 1e-7
                if method != "vb_online":
                    good = good and rel(s.train_stat, o.train_stat) < 1e-7
                if method != "mcmc":
                    good = good and rel(s.free_energy, o.free_energy) < 1e-7
                if not good:
                    print("MISMATCH", name, it, s.test_rmse, o.test_rmse, s.free_energy, o.free_energy, flush=True)
                ok = ok and good
        # replicated parameters are bit-identical on every rank
        st = E.get_state()
        digest = hashlib.sha256(np.concatenate([st["w_mean"], st["v_mean"].ravel(), st["w_var"], st["v_var"].ravel()]).tobytes()).hexdigest()
        with open(os.path.join(scratch, f"digest_{name}_{rank}.tmp"), "w") as f:
            f.write(digest)
        os.rename(os.path.join(scratch, f"digest_{name}_{rank}.tmp"), os.path.join(scratch, f"digest_{name}_{rank}"))
        if info["fused_schedule"] & 1 and method != "vb_online":
            ok = ok and E.copies_max_diff() == 0.0
        E.close()
        if rank == 0:
            for r in range(1, world):
                p = os.path.join(scratch, f"digest_{name}_{r}")
                for _ in range(60000):
                    if os.path.exists(p):
                        break
                    time.sleep(0.005)
                same = open(p).read() == digest
                if not same:
                    print("PARAMETERS DIFFER BETWEEN RANKS", name, r, flush=True)
                ok = ok and same
            print("case", name, "schedule", info["fused_schedule"], "ok" if ok else "FAILED", flush=True)
    if rank == 0:
        print("MRANK_OK" if ok else "MRANK_FAIL", flush=True)


if __name__ == "__main__":
    main()
