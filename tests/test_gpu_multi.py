"""-m gpu: the row-sharded path on >= 2 GPUs of one box (skipped on a 1-GPU box): one process per GPU, NCCL
allreduce of the per-run column sums inside the engine, results equal to the single-process oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_vb_matches_oracle(built):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", os.path.join(ROOT, "tests", "mgpu_worker.py")], capture_output=True, text=True, timeout=600)
    assert "MGPU_OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]


@pytest.mark.parametrize("shard", ["cross", "range"])
def test_cli_two_gpus_writes_the_single_gpu_files(built, tmp_path, shard):
    """bin/libFM as two processes on two GPUs (WORLD_SIZE / RANK / LOCAL_RANK / SVBFM_COMM_FILE, real NCCL): cross shards from the CLI
    (the default for two one-hot fields with x = 1) and contiguous case ranges (SVBFM_SHARD=range) write the files of the one-GPU run.
    The emulator runs the same launch with a fake NCCL in the CPU suite (tests/test_emu_kernels.py)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    G = os.path.join(ROOT, "tests", "golden")
    exe = os.path.join(ROOT, "scalable-variational-bayesian-factorization-machine_b200", "bin", "libFM")
    args = [exe, "-task", "r", "-train", os.path.join(G, "g1_train.libfm"), "-test", os.path.join(G, "g1_test.libfm"), "-dim", "1,1,4", "-method", "vb",
            "-iter", "4", "-seed", "42"]
    base = {k: v for k, v in os.environ.items() if k not in ("WORLD_SIZE", "RANK", "LOCAL_RANK", "SVBFM_COMM_FILE", "SVBFM_SHARD")}
    one = tmp_path / "one"; one.mkdir()
    r = subprocess.run(args, env=base, cwd=one, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ERROR" not in r.stderr, r.stderr[-2000:]
    want = [float(v) for v in open(one / "test_rmse_114_vb").read().split()]
    want_f = [float(v) for v in open(one / "free_energy_114_vb").read().split()]
    idfile = tmp_path / "comm_id"
    dirs, procs = [], []
    for rank in range(2):
        d = tmp_path / f"rank{rank}"; d.mkdir(); dirs.append(d)
        env = dict(base, WORLD_SIZE="2", RANK=str(rank), LOCAL_RANK=str(rank), SVBFM_COMM_FILE=str(idfile), SVBFM_COMM_NONCE=f"gpu-test-{shard}", SVBFM_SHARD=shard)
        procs.append(subprocess.Popen(args, env=env, cwd=d, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = []
    try:
        outs = [p.communicate(timeout=300) for p in procs]
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    assert all(p.returncode == 0 for p in procs) and all("ERROR" not in e for _, e in outs), [e[-1500:] for _, e in outs]
    assert ("cross shards (first field" in outs[0][0]) == (shard == "cross"), outs[0][0][-2000:]
    got = [float(v) for v in open(dirs[0] / "test_rmse_114_vb").read().split()]
    got_f = [float(v) for v in open(dirs[0] / "free_energy_114_vb").read().split()]
    assert len(got) == 4 and all(abs(a - b) <= 1e-5 * b for a, b in zip(got, want)), (got, want)
    assert all(abs(a - b) <= 1e-5 * abs(b) for a, b in zip(got_f, want_f)), (got_f, want_f)
    assert not idfile.exists()                             # rank 0 removes the record once the communicator stands
