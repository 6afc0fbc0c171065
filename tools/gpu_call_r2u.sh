#!/bin/bash
# Round 2, GPU call U (2 GPUs): the 2-GPU test at HEAD (real NCCL: case ranges, user blocks, cross shards, sharded vb_online through the
# packed batches), the CLI on two GPUs against the one-GPU files (tests/test_gpu_multi.py), the strong-scaling bench line at N = 2 (what the driver launches).
set -u
mkdir -p gpurun_out
out=gpurun_out
py=python
t0=$(date +%s)
timeout 600 $py -m pytest tests/test_gpu_multi.py -m gpu -q > $out/r2u_pytest_multi.log 2>&1; echo "pytest multi rc=$? $(tail -1 $out/r2u_pytest_multi.log) [$(( $(date +%s) - t0 )) s]" | tee $out/r2u_summary.txt
timeout 600 $py -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline \
  > $out/r2u_bench_n2.json 2> $out/r2u_bench_n2.err; echo "bench n2 rc=$? $(grep -o '"ms_per_step": [0-9.]*' $out/r2u_bench_n2.json | head -1) $(grep -o '"parity_vs_n1": {[^}]*' $out/r2u_bench_n2.json | head -c 200) [$(( $(date +%s) - t0 )) s]" | tee -a $out/r2u_summary.txt
