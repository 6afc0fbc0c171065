#!/bin/bash
# Round 2, GPU call H (8 GPUs): strong scaling with cross shards over peer memory (flags + coalesced fetch of the other ranks' blocks
# inside k_records_remote) at N = 8 and 4, against the same schedule over ncclAllGather (SVBFM_NO_P2P=1) at N = 8.
set -u
mkdir -p gpurun_out
out=gpurun_out
py=python
tr() { n=$1; port=$2; shift 2; $py -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@"; }
timeout 600 bash -c "$(declare -f tr); py=$py; tr 8 29821 --steps 5 --warmup 3 --no-e2e" > $out/r2h_bench_n8_p2p.json 2> $out/r2h_bench_n8_p2p.err; echo "bench n8 p2p rc=$? $(grep -o '"ms_per_step": [0-9.]*' $out/r2h_bench_n8_p2p.json | head -1)" | tee $out/r2h_summary.txt
SVBFM_NO_P2P=1 timeout 600 bash -c "$(declare -f tr); py=$py; tr 8 29822 --steps 5 --warmup 3 --no-e2e --no-parity" > $out/r2h_bench_n8_nccl.json 2> $out/r2h_bench_n8_nccl.err; echo "bench n8 nccl rc=$? $(grep -o '"ms_per_step": [0-9.]*' $out/r2h_bench_n8_nccl.json | head -1)" | tee -a $out/r2h_summary.txt
timeout 600 bash -c "$(declare -f tr); py=$py; tr 4 29823 --steps 5 --warmup 3 --no-e2e --no-parity" > $out/r2h_bench_n4_p2p.json 2> $out/r2h_bench_n4_p2p.err; echo "bench n4 p2p rc=$? $(grep -o '"ms_per_step": [0-9.]*' $out/r2h_bench_n4_p2p.json | head -1)" | tee -a $out/r2h_summary.txt
grep -v "^\*\*\*\|OMP_NUM" $out/r2h_bench_n8_p2p.err | tail -5
