#!/bin/bash
# First gpurun call of the next measurement round: validates and times everything that was built without a GPU
# (DESIGN.md section 7a). Every step runs under its own timeout so that a hang in an experimental kernel cannot hold the box;
# results land in gpurun_out/next_*.  Usage:
#   gpurun --timeout 3000 -- 'bash tools/next_gpu_call.sh'      (about 40 minutes of box time)
set -u
mkdir -p gpurun_out
out=gpurun_out
py=python

# 1. the default path: full GPU suite (includes the vb_online batch lists, now the default)
timeout 600 $py -m pytest tests -m gpu -x -q --deselect tests/test_gpu_zzzz_full_size.py > $out/next_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?" | tee $out/next_summary.txt

# 1b. BASELINE's full sizes (first-principles checks of the handed-back state; written without a GPU)
timeout 900 $py -m pytest tests/test_gpu_zzzz_full_size.py -m gpu -q --durations=0 > $out/next_pytest_full_size.log 2>&1; echo "full-size properties rc=$?" | tee -a $out/next_summary.txt

# 1c. the randomised differential runs on the real engine (oracle and reference binary as checkers), default and experiments
timeout 200 $py tests/fuzz_parity.py --seconds 90 --seed 101 > $out/next_fuzz_parity.log 2>&1; echo "fuzz parity rc=$?" | tee -a $out/next_summary.txt
timeout 200 $py tests/fuzz_cli.py --seconds 90 --seed 102 > $out/next_fuzz_cli.log 2>&1; echo "fuzz cli rc=$?" | tee -a $out/next_summary.txt
SVBFM_REC_RANK=1 SVBFM_GRAPH=1 timeout 200 $py tests/fuzz_parity.py --seconds 60 --seed 103 > $out/next_fuzz_parity_rank_graph.log 2>&1; echo "fuzz parity (rec_rank + graph) rc=$?" | tee -a $out/next_summary.txt

# 2. the experiments, each alone (a hang or a wrong result in one must not hide the others)
timeout 600 $py -m pytest tests/test_gpu_zz_experiments.py -m gpu -q -k 'rec_rank or graph' > $out/next_pytest_rec_rank.log 2>&1; echo "rec_rank tests rc=$?" | tee -a $out/next_summary.txt
SVBFM_RUN_EXPERIMENTS=1 timeout 900 $py -m pytest tests/test_gpu_zzz_tma_ring.py -m gpu -q > $out/next_pytest_tma.log 2>&1; echo "tma tests rc=$?" | tee -a $out/next_summary.txt

# 3. bench lines: default, each experiment, both (device-resident value only where e2e is not the question)
timeout 600 $py bench.py --steps 5 --warmup 3 > $out/next_bench_default.json 2> $out/next_bench_default.err; echo "bench default rc=$?" | tee -a $out/next_summary.txt
SVBFM_REC_RANK=1 timeout 600 $py bench.py --steps 5 --warmup 3 --no-cpu-baseline > $out/next_bench_rec_rank.json 2> $out/next_bench_rec_rank.err; echo "bench rec_rank rc=$?" | tee -a $out/next_summary.txt
SVBFM_STREAM_TMA=1 timeout 600 $py bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $out/next_bench_tma.json 2> $out/next_bench_tma.err; echo "bench tma rc=$?" | tee -a $out/next_summary.txt
SVBFM_REC_RANK=1 SVBFM_STREAM_TMA=1 timeout 600 $py bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $out/next_bench_rec_rank_tma.json 2> $out/next_bench_rec_rank_tma.err; echo "bench rec_rank+tma rc=$?" | tee -a $out/next_summary.txt
timeout 900 $py bench.py --method vb_online --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $out/next_bench_vbo.json 2> $out/next_bench_vbo.err; echo "bench vb_online rc=$?" | tee -a $out/next_summary.txt
SVBFM_VBO_FULL_PASSES=1 timeout 900 $py bench.py --method vb_online --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $out/next_bench_vbo_full_passes.json 2> $out/next_bench_vbo_full_passes.err; echo "bench vb_online (masked passes) rc=$?" | tee -a $out/next_summary.txt

for w in ml1m ml10m; do
  SVBFM_GRAPH=1 timeout 600 $py bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $out/next_bench_${w}_graph.json 2> $out/next_bench_${w}_graph.err; echo "bench $w (graph replay) rc=$?" | tee -a $out/next_summary.txt
  timeout 600 $py bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > $out/next_bench_$w.json 2> $out/next_bench_$w.err; echo "bench $w rc=$?" | tee -a $out/next_summary.txt
  SVBFM_TILE_ENTRIES=4096 timeout 600 $py bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $out/next_bench_${w}_tiles4096.json 2> $out/next_bench_${w}_tiles4096.err; echo "bench $w (4096-entry tiles) rc=$?" | tee -a $out/next_summary.txt
done
grep -h -o '"ms_per_step": [0-9.]*' $out/next_bench_*.json | paste -d' ' - - - - - - 2>/dev/null | tee -a $out/next_summary.txt
for f in $out/next_bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -1)"; done | tee -a $out/next_summary.txt

# 4. ncu (only after the plain runs above): launch list of one iteration and a full capture of two steady k_stream launches
#    (field 0 + field 1), default layout and rank-ordered records; read here with `ncu -i ... --page raw --csv`
if grep -q '"ms_per_step"' $out/next_bench_default.json 2>/dev/null; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/next_launches_default.csv \
    $py bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $out/next_ncu_launches.log 2>&1; echo "ncu launch list rc=$?" | tee -a $out/next_summary.txt
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream -s 12 -c 2 -f -o $out/next_ncu_k_stream_default \
    $py bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $out/next_ncu_default.log 2>&1; echo "ncu k_stream (default) rc=$?" | tee -a $out/next_summary.txt
  SVBFM_REC_RANK=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream -s 12 -c 2 -f -o $out/next_ncu_k_stream_rec_rank \
    $py bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $out/next_ncu_rec_rank.log 2>&1; echo "ncu k_stream (rec_rank) rc=$?" | tee -a $out/next_summary.txt
fi
