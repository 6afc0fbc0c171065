#!/bin/bash
# Round 2, GPU call N (1 GPU): ncu full capture of two batch passes of vb_online (k_stream<..., IDX>) and of k_finalize_vbo.
set -u
mkdir -p gpurun_out
out=gpurun_out
py=python
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_stream|k_finalize_vbo" -s 600 -c 4 -f -o $out/r2n_ncu_vbo \
  $py bench.py --method vb_online --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-other-configs > $out/r2n_ncu_vbo.log 2>&1; echo "ncu vbo rc=$?" | tee $out/r2n_summary.txt
