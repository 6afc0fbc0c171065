#!/bin/bash
# Round 2, GPU call E (2 GPUs): the 2-GPU parity test on real NCCL (cross shards included), the reworked bench.py at N = 1 (e2e,
# other_configs) and at N = 2 with strong scaling: cross shards (default) against user-block shards with the item allreduce.
set -u
mkdir -p gpurun_out
out=gpurun_out
py=python
timeout 900 $py -m pytest tests/test_gpu_multi.py -m gpu -x -q > $out/r2e_pytest_multi.log 2>&1; echo "pytest multi rc=$? $(tail -1 $out/r2e_pytest_multi.log)" | tee $out/r2e_summary.txt
timeout 900 $py bench.py --steps 5 --warmup 3 --no-cpu-baseline > $out/r2e_bench_n1.json 2> $out/r2e_bench_n1.err; echo "bench n1 rc=$? $(grep -o '"ms_per_step": [0-9.]*' $out/r2e_bench_n1.json | head -1)" | tee -a $out/r2e_summary.txt
tr() { $py -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 "${@:2}"; }
timeout 900 bash -c "$(declare -f tr); py=$py; tr 29701 --steps 5 --warmup 3" > $out/r2e_bench_n2_cross.json 2> $out/r2e_bench_n2_cross.err; echo "bench n2 cross rc=$? $(grep -o '"ms_per_step": [0-9.]*' $out/r2e_bench_n2_cross.json | head -1)" | tee -a $out/r2e_summary.txt
timeout 900 bash -c "$(declare -f tr); py=$py; tr 29702 --steps 5 --warmup 3 --shard-by user_block --no-e2e" > $out/r2e_bench_n2_userblock.json 2> $out/r2e_bench_n2_userblock.err; echo "bench n2 user_block rc=$? $(grep -o '"ms_per_step": [0-9.]*' $out/r2e_bench_n2_userblock.json | head -1)" | tee -a $out/r2e_summary.txt
timeout 600 bash -c "$(declare -f tr); py=$py; tr 29703 --steps 3 --warmup 2 --method mcmc --no-e2e" > $out/r2e_bench_n2_cross_mcmc.json 2> $out/r2e_bench_n2_cross_mcmc.err; echo "bench n2 cross mcmc rc=$? $(grep -o '"ms_per_step": [0-9.]*' $out/r2e_bench_n2_cross_mcmc.json | head -1)" | tee -a $out/r2e_summary.txt
tail -5 $out/r2e_bench_n2_cross.err
