#!/bin/bash
# Round 2, GPU call I (1 GPU): the whole -m gpu suite at HEAD (cross shards, model save / load, device transpose + svbfm_set_csr, the
# reference's own data as a CLI golden, ML-1M oracle parity, MCMC hyper trajectories), then the default bench line.
set -u
mkdir -p gpurun_out
out=gpurun_out
py=python
timeout 1500 $py -m pytest tests -m gpu -x -q > $out/r2i_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$? $(tail -1 $out/r2i_pytest_gpu.log)" | tee $out/r2i_summary.txt
timeout 900 $py bench.py > $out/r2i_bench_default.json 2> $out/r2i_bench_default.err; echo "bench default rc=$? $(grep -o '"ms_per_step": [0-9.]*' $out/r2i_bench_default.json | head -1)" | tee -a $out/r2i_summary.txt
tail -5 $out/r2i_bench_default.err
