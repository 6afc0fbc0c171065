"""Kernel logic on a machine without a GPU: a subset of the GPU parity tests (tests/test_gpu_parity.py, unchanged) run in a
subprocess against tests/emu/_build/libsvbfm_emu.so -- csrc/*.cu compiled for the host on a lockstep CUDA emulator (every
CUDA thread a fiber, warp primitives and __syncthreads as barriers; tests/emu/include/cuda_runtime.h). This checks index
arithmetic, warp-synchronous code and the engine's launch order against the oracle; it says nothing about the B200 (speed,
memory system, races between CTAs): the `-m gpu` tests do. The package itself never loads the emulated library."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# quick ones that between them reach k_stream (VB, MCMC, small tiles, x != 1), the general per-run schedule, k_predict /
# k_predict2, the finalize / hyper-parameter kernels and the device ingest
SUBSET = ["test_vb_two_field_onehot", "test_vb_ragged_multihot", "test_vb_groups_and_small_tiles", "test_mcmc_als_exact",
          "test_vb_values_no_reorder", "test_stream_schedule_sorted_input", "test_vb_k_zero_and_empty_test_columns",
          "test_vb_online_vs_oracle"]      # vb_online: packed batches, k_stream_rows, dense column ids against the oracle


@pytest.fixture(scope="module")
def emu_lib(built):
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    return build_emu.build()


def test_emulated_library_exports_the_abi(emu_lib):
    import ctypes
    import svbfm_b200 as sv
    L = ctypes.CDLL(emu_lib)
    for name in sv.ABI_SYMBOLS:
        assert hasattr(L, name), name


def test_parity_subset_on_the_emulator(emu_lib):
    env = dict(os.environ, SVBFM_LIB=emu_lib)
    cmd = [sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-x", "-q", "-p", "no:cacheprovider",
           "-k", " or ".join(SUBSET)]
    r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert "passed" in r.stdout


def test_two_ranks_on_the_emulator(emu_lib, tmp_path):
    """The sharded launch order (allreduce of the column sums, exclusive user blocks + block exchange, agreement on the schedule,
    sharded vb_online) with two PROCESSES: the host build of the kernels + tests/emu/fake_nccl.c (the nine NCCL entry points the
    engine binds, over shared memory). Same cases as the 2-GPU test (tests/mgpu_worker.py); rank 0 compares with the oracle."""
    build_dir = os.path.dirname(emu_lib)
    env = dict(os.environ, SVBFM_LIB=emu_lib, LD_LIBRARY_PATH=build_dir + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    worker = os.path.join(ROOT, "tests", "emu", "mrank_worker.py")
    procs = [subprocess.Popen([sys.executable, worker, str(r), "2", str(tmp_path)], env=env, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = []
    for p in procs:
        try:
            outs.append(p.communicate(timeout=900)[0])
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    assert all(p.returncode == 0 for p in procs), "\n".join(o[-2000:] for o in outs)
    assert "MRANK_OK" in outs[0], outs[0][-3000:]


def test_experimental_variants_on_the_emulator(emu_lib):
    """One case each of the kernel variants behind knobs: the bulk-copy ring of k_stream against plain loads (bit-identical; misaligned
    second field, partial last batch), the records in column order (SVBFM_REC_RANK=0), and the vb_online batch passes four ways (index
    lists, packed through k_stream, packed through k_stream_rows with dense and with global column ids)."""
    env = dict(os.environ, SVBFM_LIB=emu_lib)
    cmd = [sys.executable, "-m", "pytest", "-m", "gpu", "-x", "-q", "-p", "no:cacheprovider",
           os.path.join(ROOT, "tests", "test_gpu_zzz_tma_ring.py") + "::test_stream_tma_ring[256-20002]",
           os.path.join(ROOT, "tests", "test_gpu_zz_experiments.py") + "::test_rec_rank_layout[0-64]",
           os.path.join(ROOT, "tests", "test_gpu_zz_experiments.py") + "::test_vb_online_packed_batches_equal_index_lists[True-64]",
           os.path.join(ROOT, "tests", "test_gpu_zy_errors.py")]           # and the error convention of the C-ABI
    r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "5 passed" in r.stdout, r.stdout[-4000:] + r.stderr[-2000:]


def test_no_access_outside_the_device_buffers(emu_lib):
    """SVBFM_EMU_GUARD=1: the emulator ends every device allocation in front of an inaccessible page (and the block cache of the emulated
    build keeps the caller's sizes), so a kernel that reads or writes past a buffer dies with SIGSEGV at the access. compute-sanitizer is
    not available on the GPU pool; this is the bounds check of the index arithmetic. VB and als on ragged / two-field data with small
    tiles, and the vb_online batch passes four ways (k_vbo_pack, k_stream_rows, dense column ids, k_finalize_vbo's long spans)."""
    env = dict(os.environ, SVBFM_LIB=emu_lib, SVBFM_EMU_GUARD="1")
    t = os.path.join(ROOT, "tests", "test_gpu_parity.py")
    cmd = [sys.executable, "-m", "pytest", "-m", "gpu", "-x", "-q", "-p", "no:cacheprovider",
           t + "::test_vb_groups_and_small_tiles", t + "::test_mcmc_als_exact", t + "::test_vb_ragged_multihot", t + "::test_vb_online_long_spans_in_a_batch",
           os.path.join(ROOT, "tests", "test_gpu_zz_experiments.py") + "::test_vb_online_packed_batches_equal_index_lists[False-32]"]
    r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "5 passed" in r.stdout, r.stdout[-4000:] + r.stderr[-2000:]


def test_full_size_properties_on_divided_shapes(emu_lib):
    """tests/test_gpu_zzzz_full_size.py (first-principles checks of the handed-back state at BASELINE's full sizes) on shapes divided
    down to a few thousand ratings: proves the test's own arithmetic (torch fp64 restatement of y-hat, T, the hyper-parameters and the
    free energy) against the engine before it meets 200 M ratings on a B200."""
    env = dict(os.environ, SVBFM_LIB=emu_lib)
    f = os.path.join(ROOT, "tests", "test_gpu_zzzz_full_size.py")
    cmd = [sys.executable, "-m", "pytest", "-m", "gpu", "-x", "-q", "-p", "no:cacheprovider",
           f + "::test_vb_full_size_state_is_consistent[ml1m]", f + "::test_mcmc_full_size_state_is_consistent[1]"]
    r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "2 passed" in r.stdout, r.stdout[-4000:] + r.stderr[-2000:]


def test_randomised_differential_run(emu_lib):
    """tests/fuzz_parity.py: random data shapes (one-hot, real values, ragged, three fields), methods, switches, groups and tile
    sizes against the oracle, statistics of every iteration and the final parameters; 60 cases here, thousands when run by hand."""
    env = dict(os.environ, SVBFM_LIB=emu_lib)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fuzz_parity.py"), "--cases", "60", "--seed", "5", "--seconds", "600"],
                       env=env, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "60 cases, no mismatch" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


def test_randomised_sharded_run(emu_lib):
    """tests/fuzz_sharded.py: three PROCESSES (fake NCCL), random shapes / methods / switches / shard modes (contiguous case ranges,
    user blocks, ranks without a case) against the single-process oracle; parameters bit-identical on every rank. 40 cases here."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fuzz_sharded.py"), "--world", "3", "--cases", "40", "--seed", "7", "--seconds", "600"],
                       cwd=ROOT, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0 and "40 cases, no mismatch" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


def test_randomised_cli_against_the_reference_binary(emu_lib):
    """tests/fuzz_cli.py: random small data sets and command lines through the UNMODIFIED reference binary (oracle/_ref/libFM) and
    through bin/libFM (engine = the emulated build), files in the CWD and `Train=` lines compared at 1e-4. 30 cases here."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libFM")):
        pytest.skip("oracle/_ref/libFM not built")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "fuzz_cli.py"), "--cases", "30", "--seed", "9", "--seconds", "600"],
                       env=dict(os.environ, SVBFM_EMU="1"), cwd=ROOT, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0 and "30 cases, no mismatch" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


def test_reference_tree_binding_on_the_emulator(emu_lib):
    """oracle/_ref/libFM_cuda (the reference's own libfm.cpp + host/reference_tree/*.h, recipe oracle/make_ref_cuda.py) with the
    emulated engine preloaded: `-method vb_cuda` against `-method vb` and `als_cuda` against `als` from the one binary."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libFM_cuda")):
        pytest.skip("oracle/_ref/libFM_cuda not built")
    f = os.path.join(ROOT, "tests", "test_gpu_ref_tree_binding.py")
    cmd = [sys.executable, "-m", "pytest", "-m", "gpu", "-x", "-q", "-p", "no:cacheprovider",
           f + "::test_vb_cuda_equals_vb_in_the_reference_binary[g1-1,1,4]", f + "::test_als_cuda_equals_als_in_the_reference_binary"]
    r = subprocess.run(cmd, env=dict(os.environ, SVBFM_EMU_PRELOAD=emu_lib), cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "2 passed" in r.stdout, r.stdout[-4000:] + r.stderr[-2000:]


def test_round2_cli_paths_on_the_emulator(emu_lib):
    """The CLI-level additions of round 2 with the emulated engine preloaded into bin/libFM and bin/transpose: -save_model / -load_model
    (bit-for-bit resume), `transpose --device` against the reference tool's bytes, rows-only binary input through svbfm_set_csr."""
    f = os.path.join(ROOT, "tests", "test_gpu_golden_cli.py")
    cmd = [sys.executable, "-m", "pytest", "-m", "gpu", "-x", "-q", "-p", "no:cacheprovider",
           f + "::test_cli_save_and_load_model_resume_bit_for_bit", f + "::test_transpose_tool_on_the_device_writes_the_reference_bytes",
           f + "::test_cli_rows_only_binary_input_is_transposed_on_the_device"]
    r = subprocess.run(cmd, env=dict(os.environ, SVBFM_LIB=emu_lib, LD_PRELOAD=emu_lib), cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "4 passed" in r.stdout, r.stdout[-4000:] + r.stderr[-2000:]


def test_device_transpose_and_row_wise_ingest_on_the_emulator(emu_lib):
    """svbfm_transpose_csr against the host transpose (one-hot and ragged multi-hot data with real values) and svbfm_set_csr against
    svbfm_set_csc: identical statistics."""
    code = r'''
import sys, os
sys.path[:0] = [ROOT, ROOT + "/oracle", ROOT + "/tests"]
import numpy as np
import svbfm_b200 as sv
from helpers import two_field, ragged, to_csc
for tr in (two_field(5000, 10, 120, 90, seed=3)[0], ragged(3000, 10, 60, seed=4)[0]):
    csc = to_csc(tr)
    cp, ci, xt = sv.transpose_csr(tr.rowptr, tr.col, tr.val, tr.n_feat)
    assert np.array_equal(cp, csc.colptr) and np.array_equal(ci, csc.case_id) and np.array_equal(xt, csc.x)
try:
    sv.transpose_csr(np.array([0, 1], dtype=np.uint64), np.array([7], dtype=np.uint32), np.ones(1, dtype=np.float32), 3)
    raise SystemExit("a feature id beyond num_cols must be refused")
except sv.SvbfmError as ex:
    assert "out of range" in str(ex)
for (tr, te), K in ((two_field(6000, 600, 100, 80, seed=5), 3), (ragged(2000, 300, 40, seed=6), 2)):
    D = max(tr.n_feat, te.n_feat) + 1
    out = []
    for mode in ("csc", "csr"):
        E = sv.Engine("vb", D, K, 1, 1, float(tr.y.min()), float(tr.y.max()), seed=42)
        if mode == "csc":
            E.set_csc(sv.TRAIN, to_csc(tr)); E.set_csc(sv.TEST, to_csc(te))
        else:
            E.set_csr(sv.TRAIN, tr.rowptr, tr.col, tr.val, tr.y, tr.n_feat); E.set_csr(sv.TEST, te.rowptr, te.col, te.val, te.y, te.n_feat)
        E.set_state(sv.host_init_state(42, D, K, 0.1, sv.VB)); E.begin()
        out.append([(s.test_rmse, s.free_energy, s.alpha) for s in E.run(3)])
        E.close()
    assert out[0] == out[1], out
print("ROW_WISE_OK")
'''.replace("ROOT", repr(ROOT))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, SVBFM_LIB=emu_lib), cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ROW_WISE_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("shard", ["cross", "range"])
def test_cli_two_ranks_on_the_emulator(emu_lib, tmp_path, shard):
    """bin/libFM as two PROCESSES (WORLD_SIZE / RANK / SVBFM_COMM_FILE, fake NCCL, emulated engine preloaded): the id record with its
    launch nonce, cross shards from the CLI (the default for two one-hot fields) and contiguous case ranges (SVBFM_SHARD=range) write
    the files of the single-process run; a stale record of an earlier launch in the same path is ignored and removed."""
    G = os.path.join(ROOT, "tests", "golden")
    exe = os.path.join(ROOT, "scalable-variational-bayesian-factorization-machine_b200", "bin", "libFM")
    build_dir = os.path.dirname(emu_lib)
    args = [exe, "-task", "r", "-train", os.path.join(G, "g1_train.libfm"), "-test", os.path.join(G, "g1_test.libfm"), "-dim", "1,1,4", "-method", "vb",
            "-iter", "4", "-seed", "42"]
    base = dict(os.environ, SVBFM_LIB=emu_lib, LD_PRELOAD=emu_lib, LD_LIBRARY_PATH=build_dir + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    one = tmp_path / "one"; one.mkdir()
    r = subprocess.run(args, env=base, cwd=one, capture_output=True, text=True, timeout=600)
    assert "ERROR" not in r.stderr, r.stderr
    want = [float(v) for v in open(one / "test_rmse_114_vb").read().split()]
    want_f = [float(v) for v in open(one / "free_energy_114_vb").read().split()]
    idfile = tmp_path / "comm_id"
    idfile.write_bytes(b"SVBFMID1" + bytes(136))          # a record of some earlier launch: wrong nonce
    dirs, procs = [], []
    for rank in range(2):
        d = tmp_path / f"rank{rank}"; d.mkdir(); dirs.append(d)
        env = dict(base, WORLD_SIZE="2", RANK=str(rank), LOCAL_RANK="0", SVBFM_COMM_FILE=str(idfile), SVBFM_COMM_NONCE="test-launch", SVBFM_SHARD=shard)
        procs.append(subprocess.Popen(args, env=env, cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=600) for p in procs]
    assert all("ERROR" not in e for _, e in outs), [e for _, e in outs]
    assert ("cross shards (first field" in outs[0][0]) == (shard == "cross"), outs[0][0][-2000:]
    got = [float(v) for v in open(dirs[0] / "test_rmse_114_vb").read().split()]
    got_f = [float(v) for v in open(dirs[0] / "free_energy_114_vb").read().split()]
    assert len(got) == 4 and all(abs(a - b) <= 1e-5 * b for a, b in zip(got, want)), (got, want)
    assert all(abs(a - b) <= 1e-5 * abs(b) for a, b in zip(got_f, want_f)), (got_f, want_f)
    assert not idfile.exists()                             # rank 0 removes the record once the communicator stands
