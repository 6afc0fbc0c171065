"""Randomised differential run of the SHARDED paths against the single-process oracle (test infrastructure; not part of the
product): W processes, one shard of the cases each, collectives through NCCL -- on the emulator through tests/emu/fake_nccl.c.
Every rank draws the same list of cases from the seed: two one-hot fields (with or without real values) or ragged multi-hot
data; vb / als / vb_online; K, k0, k1, tile size; contiguous case ranges or user blocks; more ranks than cases. Rank 0 compares
every iteration's statistics with the oracle; the replicated parameters must be bit-identical on every rank.

  python tests/fuzz_sharded.py --world 3 --seconds 300 --seed 1        # launcher: builds / finds the emulator, spawns the ranks
"""
import argparse
import hashlib
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]


def exchange(scratch, tag, rank, payload=None):
    path = os.path.join(scratch, tag)
    if rank == 0:
        with open(path + ".tmp", "wb") as f:
            f.write(payload)
        os.rename(path + ".tmp", path)
        return payload
    for _ in range(120000):
        if os.path.exists(path):
            return open(path, "rb").read()
        time.sleep(0.005)
    raise RuntimeError("rank 0 never published " + tag)


def worker(a):
    import numpy as np
    import oracle_binding as ob
    import svbfm_b200 as sv
    from helpers import ragged, rel, to_csc, two_field
    d = sv.submodule("dist")
    rank, world, scratch = a.rank, a.world, a.scratch
    r = np.random.default_rng(a.seed)
    t0, n_done, ok = time.time(), 0, True
    while n_done < a.cases:
        # rank 0 decides whether there is time for another case (the ranks' clocks differ)
        go = exchange(scratch, f"go_{n_done}", rank, b"1" if time.time() - t0 < a.seconds else b"0")
        if go != b"1":
            break
        kind = r.choice(["two", "two", "two_values", "ragged"])
        method = r.choice(["vb", "als", "vb_online", "vb_online"])
        N = int(r.choice([1, 5, 40, 600, 5000]))
        K = int(r.choice([0, 1, 3]))
        k0, k1 = int(r.random() < 0.8), int(r.random() < 0.8)
        tile = int(r.choice([0, 32, 64]))
        seed = int(r.integers(1, 10_000))
        U, I = int(r.choice([3, 20, 150])), int(r.choice([2, 15, 120]))
        blocks = bool(r.random() < 0.5) and kind != "ragged"
        cross = blocks and bool(r.random() < 0.6)       # cross shards: the second residual copy on a shard by item block (may be refused: every rank alike)
        nb = int(r.choice([1, 3, 8]))
        iters = int(r.choice([1, 2, 3]))
        if kind == "ragged":
            tr, te = ragged(N, 50, int(r.choice([6, 40])), seed=seed, max_nnz=int(r.choice([2, 5])))
        else:
            tr, te = two_field(N, 50, U, I, seed=seed, values=(kind == "two_values"))
        if method == "vb_online" and -(-N // nb) * (nb - 1) >= N:
            nb = 1
        m = "mcmc" if method == "als" else method
        D = max(tr.n_feat, te.n_feat) + (0 if m == "vb_online" else 1)
        desc = f"case {n_done}: {kind} {method} N={N} K={K} k0={k0} k1={k1} tile={tile} seed={seed} U={U} I={I} blocks={blocks} cross={cross} nb={nb} iters={iters}"
        n_done += 1
        if D == 0 or (m != "vb_online" and D == 1):
            continue
        kw = dict(do_sample=False, do_multilevel=False) if m == "mcmc" else {}
        okw = dict(kw, num_batch=nb) if m == "vb_online" else kw
        buf = (sv.C.c_uint8 * sv.COMM_ID_BYTES)()
        if rank == 0:
            assert sv.lib().svbfm_comm_get_unique_id(buf) == 0
        uid = exchange(scratch, f"uid_{n_done}", rank, bytes(buf))
        E = sv.Engine(m, D, K, k0, k1, float(tr.y.min()), float(tr.y.max()), seed=42, tile_entries=tile, **kw)
        E.comm_init(uid, rank, world)
        if blocks:
            shard, mine = d.shard_csc_by_block(to_csc(tr), rank, world, U)
        else:
            shard = d.shard_csc(to_csc(tr), rank, world)
            lo, hi = d.shard_bounds(tr.n_rows, rank, world)
            mine = np.arange(lo, hi)
        E.set_csc(sv.TRAIN, shard)
        if cross:
            try:
                E.set_csc(sv.TRAIN_SECOND, d.shard_csc_by_second_block(to_csc(tr), rank, world, U)[0])
            except sv.SvbfmError as ex:      # data that does not qualify (x != 1, vb_online, no stream schedule): the handle carries on without it
                assert "TRAIN_SECOND" in str(ex), ex
        E.set_csc(sv.TEST, d.shard_csc(to_csc(te), rank, world))
        info = E.info()
        E.set_state(sv.host_init_state(42, D, K, 0.1, sv.METHODS[m]))
        E.begin()
        if m == "vb_online":
            n = tr.n_rows
            size_except_last = int(np.ceil(n / nb))
            shuffle = np.arange(1, n + 1, dtype=np.uint32)
            hist = []
            for _ in range(iters):
                sv.lib().svbfm_host_random_shuffle(shuffle.ctypes.data_as(sv.C.c_void_p), n)
                batch = (np.ceil(shuffle.astype(np.float64) / size_except_last) - 1).astype(np.uint32)
                hist.append(E.vb_online_epoch(np.ascontiguousarray(batch[mine]), nb))
        else:
            hist = E.run(iters)
        good = True
        stol = 1e-7 if kind == "two" else 2e-6          # real values: DESIGN.md section 2 (float-rounding residue of the reference's caches)
        same = lambda x, y: (np.isnan(x) and np.isnan(y)) or (abs(x) < 1e-8 and abs(y) < 1e-8) or rel(x, y) < stol
        if rank == 0:
            orc = ob.Oracle(m, tr, te, K=K, seed=42, k0=k0, k1=k1, **okw)
            for it, s in enumerate(hist):
                o = orc.iterate()
                names = ("test_rmse",) + (("train_stat",) if m != "vb_online" else ()) + (("free_energy", "alpha") if m != "mcmc" else ())
                for name in names:
                    if not same(getattr(s, name), getattr(o, name)):
                        print(desc, f"MISMATCH iteration {it}: {name} {getattr(s, name)!r} != {getattr(o, name)!r}", flush=True)
                        good = False
        st = E.get_state()
        digest = hashlib.sha256(np.concatenate([st["w_mean"], st["v_mean"].ravel(), st["w_var"], st["v_var"].ravel()]).tobytes()).hexdigest().encode()
        if rank:
            with open(os.path.join(scratch, f"digest_{n_done}_{rank}.tmp"), "wb") as f:
                f.write(digest)
            os.rename(os.path.join(scratch, f"digest_{n_done}_{rank}.tmp"), os.path.join(scratch, f"digest_{n_done}_{rank}"))
        if info["fused_schedule"] & 1 and m != "vb_online" and E.copies_max_diff() != 0.0:
            print(desc, "RESIDUAL COPIES DIFFER on rank", rank, flush=True)
            good = False
        E.close()
        if rank == 0:
            for q in range(1, world):
                p = os.path.join(scratch, f"digest_{n_done}_{q}")
                for _ in range(120000):
                    if os.path.exists(p):
                        break
                    time.sleep(0.005)
                if open(p, "rb").read() != digest:
                    print(desc, "PARAMETERS DIFFER BETWEEN RANKS", q, flush=True)
                    good = False
            print(desc, f"schedule={info['fused_schedule']} excl={info['exclusive_blocks']}", "OK" if good else "FAILED", flush=True)
        ok = ok and good
        if not ok:
            break
    if rank == 0:
        print(f"{n_done} cases, no mismatch" if ok else "SHARDED_FUZZ_FAILED", flush=True)
    sys.exit(0 if ok else 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=2)
    ap.add_argument("--seconds", type=float, default=120)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cases", type=int, default=10**9)
    ap.add_argument("--rank", type=int, default=-1)
    ap.add_argument("--scratch", type=str, default="")
    a = ap.parse_args()
    if a.rank >= 0:
        return worker(a)
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    lib = build_emu.build()
    bdir = os.path.dirname(lib)
    env = dict(os.environ, SVBFM_LIB=lib, LD_LIBRARY_PATH=bdir + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    with tempfile.TemporaryDirectory() as scratch:
        procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--rank", str(q), "--world", str(a.world), "--scratch", scratch, "--seed", str(a.seed),
                                   "--seconds", str(a.seconds), "--cases", str(a.cases)], env=env, cwd=ROOT) for q in range(a.world)]
        deadline = time.time() + a.seconds + 600
        rc = 0
        while any(p.poll() is None for p in procs):
            if time.time() > deadline or any(p.poll() not in (None, 0) for p in procs):
                time.sleep(2)
                for p in procs:
                    if p.poll() is None:
                        p.kill()
                rc = 1
                break
            time.sleep(0.2)
        rc = rc or max((p.returncode or 0) for p in procs)
    sys.exit(1 if rc else 0)


if __name__ == "__main__":
    main()
