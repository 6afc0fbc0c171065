#!/bin/bash
# Round 2, GPU call A: time the variants built without a GPU at the end of round 1 (rank-ordered records, TMA ring, graph replay,
# vb_online batch lists). Every step under its own timeout; results in gpurun_out/r2a_*.
set -u
mkdir -p gpurun_out
out=gpurun_out
py=python
B="--no-cpu-baseline --no-e2e"
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $out/r2a_gpu.txt 2>&1
SVBFM_RUN_EXPERIMENTS=1 timeout 300 $py -m pytest tests/test_gpu_zzz_tma_ring.py tests/test_gpu_zz_experiments.py -m gpu -q -x > $out/r2a_pytest_experiments.log 2>&1; echo "experiments tests rc=$?" | tee $out/r2a_summary.txt
run() { name=$1; shift; timeout 400 env "$@" $py bench.py --steps 5 --warmup 3 $B > $out/r2a_bench_$name.json 2> $out/r2a_bench_$name.err; echo "bench $name rc=$? $(grep -o '"ms_per_step": [0-9.]*' $out/r2a_bench_$name.json | head -1)" | tee -a $out/r2a_summary.txt; }
run default SVBFM_X=0
run rec_rank SVBFM_REC_RANK=1
run tma SVBFM_STREAM_TMA=1
run rec_rank_tma SVBFM_REC_RANK=1 SVBFM_STREAM_TMA=1
for w in ml1m ml10m; do
  timeout 300 $py bench.py --workload $w --steps 10 --warmup 3 $B > $out/r2a_bench_$w.json 2> $out/r2a_bench_$w.err; echo "bench $w rc=$? $(grep -o '"ms_per_step": [0-9.]*' $out/r2a_bench_$w.json | head -1)" | tee -a $out/r2a_summary.txt
  SVBFM_GRAPH=1 timeout 300 $py bench.py --workload $w --steps 10 --warmup 3 $B > $out/r2a_bench_${w}_graph.json 2> $out/r2a_bench_${w}_graph.err; echo "bench $w graph rc=$? $(grep -o '"ms_per_step": [0-9.]*' $out/r2a_bench_${w}_graph.json | head -1)" | tee -a $out/r2a_summary.txt
done
timeout 400 $py bench.py --method vb_online --steps 2 --warmup 1 $B > $out/r2a_bench_vbo.json 2> $out/r2a_bench_vbo.err; echo "bench vbo rc=$? $(grep -o '"ms_per_step": [0-9.]*' $out/r2a_bench_vbo.json | head -1)" | tee -a $out/r2a_summary.txt
timeout 400 $py bench.py --method mcmc --steps 3 --warmup 2 $B > $out/r2a_bench_mcmc.json 2> $out/r2a_bench_mcmc.err; echo "bench mcmc rc=$? $(grep -o '"ms_per_step": [0-9.]*' $out/r2a_bench_mcmc.json | head -1)" | tee -a $out/r2a_summary.txt
# ncu: full capture of two steady k_stream launches (field 0 + field 1), default and rank-ordered records
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stream -s 12 -c 2 -f -o $out/r2a_ncu_k_stream_default \
  $py bench.py --steps 1 --warmup 1 $B > $out/r2a_ncu_default.log 2>&1; echo "ncu default rc=$?" | tee -a $out/r2a_summary.txt
SVBFM_REC_RANK=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stream -s 12 -c 2 -f -o $out/r2a_ncu_k_stream_rec_rank \
  $py bench.py --steps 1 --warmup 1 $B > $out/r2a_ncu_rec_rank.log 2>&1; echo "ncu rec_rank rc=$?" | tee -a $out/r2a_summary.txt
SVBFM_STREAM_TMA=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stream -s 12 -c 2 -f -o $out/r2a_ncu_k_stream_tma \
  $py bench.py --steps 1 --warmup 1 $B > $out/r2a_ncu_tma.log 2>&1; echo "ncu tma rc=$?" | tee -a $out/r2a_summary.txt
