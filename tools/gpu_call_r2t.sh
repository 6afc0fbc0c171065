#!/bin/bash
# Round 2, GPU call T (1 GPU): the whole -m gpu suite at HEAD (no -x: every failure listed), the default bench line (what the driver
# runs), the reference arm, the ncu launch list of the bench command.
set -u
mkdir -p gpurun_out
out=gpurun_out
py=python
t0=$(date +%s)
timeout 1500 $py -m pytest tests -m gpu -q > $out/r2t_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$? $(tail -1 $out/r2t_pytest_gpu.log) [$(( $(date +%s) - t0 )) s]" | tee $out/r2t_summary.txt
timeout 900 $py bench.py > $out/r2t_bench_default.json 2> $out/r2t_bench_default.err; echo "bench default rc=$? $(grep -o '"ms_per_step": [0-9.]*' $out/r2t_bench_default.json | head -1) [$(( $(date +%s) - t0 )) s]" | tee -a $out/r2t_summary.txt
timeout 600 $py bench.py --impl reference --steps 2 --warmup 1 > $out/r2t_bench_reference.json 2> $out/r2t_bench_reference.err; echo "bench reference rc=$? [$(( $(date +%s) - t0 )) s]" | tee -a $out/r2t_summary.txt
timeout 300 $py -c 'import __graft_entry__ as g; g.smoke()' > $out/r2t_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 $out/r2t_smoke.log) [$(( $(date +%s) - t0 )) s]" | tee -a $out/r2t_summary.txt
