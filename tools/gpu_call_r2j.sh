#!/bin/bash
# Round 2, GPU call J (8 GPUs): cross shards with the own block PUSHED into every rank's stage (coalesced stores, k_push_block) at N = 8 and 2.
set -u
mkdir -p gpurun_out
out=gpurun_out
py=python
tr() { n=$1; port=$2; shift 2; $py -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@"; }
timeout 500 bash -c "$(declare -f tr); py=$py; tr 8 29831 --steps 5 --warmup 3 --no-e2e" > $out/r2j_bench_n8_push.json 2> $out/r2j_bench_n8_push.err; echo "bench n8 push rc=$? $(grep -o '"ms_per_step": [0-9.]*' $out/r2j_bench_n8_push.json | head -1)" | tee $out/r2j_summary.txt
timeout 500 bash -c "$(declare -f tr); py=$py; tr 2 29832 --steps 5 --warmup 3 --no-e2e --no-parity" > $out/r2j_bench_n2_push.json 2> $out/r2j_bench_n2_push.err; echo "bench n2 push rc=$? $(grep -o '"ms_per_step": [0-9.]*' $out/r2j_bench_n2_push.json | head -1)" | tee -a $out/r2j_summary.txt
grep -v "^\*\*\*\|OMP_NUM" $out/r2j_bench_n8_push.err | tail -5
