#!/bin/bash
# Round 2, GPU call G (2 GPUs): the peer-memory exchange of the cross shards on real NVLink (k_finalize stores {mean, var} into every
# rank's stage over CUDA IPC mappings, k_records_remote exchanges flag words): 2-GPU parity test, bench at N = 2 with and without it.
set -u
mkdir -p gpurun_out
out=gpurun_out
py=python
timeout 900 $py -m pytest tests/test_gpu_multi.py -m gpu -x -q > $out/r2g_pytest_multi.log 2>&1; echo "pytest multi (p2p) rc=$? $(tail -1 $out/r2g_pytest_multi.log)" | tee $out/r2g_summary.txt
SVBFM_NO_P2P=1 timeout 900 $py -m pytest tests/test_gpu_multi.py -m gpu -x -q > $out/r2g_pytest_multi_nccl.log 2>&1; echo "pytest multi (nccl) rc=$? $(tail -1 $out/r2g_pytest_multi_nccl.log)" | tee -a $out/r2g_summary.txt
tr() { n=$1; port=$2; shift 2; $py -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@"; }
timeout 600 bash -c "$(declare -f tr); py=$py; tr 2 29811 --steps 5 --warmup 3 --no-e2e" > $out/r2g_bench_n2_p2p.json 2> $out/r2g_bench_n2_p2p.err; echo "bench n2 p2p rc=$? $(grep -o '"ms_per_step": [0-9.]*' $out/r2g_bench_n2_p2p.json | head -1)" | tee -a $out/r2g_summary.txt
SVBFM_NO_P2P=1 timeout 600 bash -c "$(declare -f tr); py=$py; tr 2 29812 --steps 5 --warmup 3 --no-e2e --no-parity" > $out/r2g_bench_n2_nccl.json 2> $out/r2g_bench_n2_nccl.err; echo "bench n2 nccl rc=$? $(grep -o '"ms_per_step": [0-9.]*' $out/r2g_bench_n2_nccl.json | head -1)" | tee -a $out/r2g_summary.txt
grep -v "^\*\*\*\|OMP_NUM" $out/r2g_bench_n2_p2p.err | tail -5
