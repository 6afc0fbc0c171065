#!/bin/bash
# Round 2, GPU call F (8 GPUs): strong scaling of the 200 M x K=50 VB iteration: cross shards (16 B per updated column allgathered,
# records formed locally) at N = 8 and 4 against user-block shards with the item allreduce at N = 8.
set -u
mkdir -p gpurun_out
out=gpurun_out
py=python
tr() { n=$1; port=$2; shift 2; $py -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@"; }
timeout 600 bash -c "$(declare -f tr); py=$py; tr 8 29801 --steps 5 --warmup 3 --no-e2e" > $out/r2f_bench_n8_cross.json 2> $out/r2f_bench_n8_cross.err; echo "bench n8 cross rc=$? $(grep -o '"ms_per_step": [0-9.]*' $out/r2f_bench_n8_cross.json | head -1)" | tee $out/r2f_summary.txt
timeout 600 bash -c "$(declare -f tr); py=$py; tr 8 29802 --steps 5 --warmup 3 --shard-by user_block --no-e2e --no-parity" > $out/r2f_bench_n8_userblock.json 2> $out/r2f_bench_n8_userblock.err; echo "bench n8 user_block rc=$? $(grep -o '"ms_per_step": [0-9.]*' $out/r2f_bench_n8_userblock.json | head -1)" | tee -a $out/r2f_summary.txt
timeout 600 bash -c "$(declare -f tr); py=$py; tr 4 29803 --steps 5 --warmup 3 --no-e2e --no-parity" > $out/r2f_bench_n4_cross.json 2> $out/r2f_bench_n4_cross.err; echo "bench n4 cross rc=$? $(grep -o '"ms_per_step": [0-9.]*' $out/r2f_bench_n4_cross.json | head -1)" | tee -a $out/r2f_summary.txt
timeout 600 bash -c "$(declare -f tr); py=$py; tr 2 29804 --steps 5 --warmup 3 --no-e2e --no-parity" > $out/r2f_bench_n2_cross.json 2> $out/r2f_bench_n2_cross.err; echo "bench n2 cross rc=$? $(grep -o '"ms_per_step": [0-9.]*' $out/r2f_bench_n2_cross.json | head -1)" | tee -a $out/r2f_summary.txt
tail -3 $out/r2f_bench_n8_cross.err
