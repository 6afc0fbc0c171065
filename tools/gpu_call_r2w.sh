#!/bin/bash
# Round 2, GPU call W (1 GPU): ncu --set full at HEAD: the two steady field passes of k_stream (roofline_traffic.json), and a batch pass +
# finalize of vb_online with dense column ids (k_stream_rows, k_finalize_vbo) on 2 M-entry batches.
set -u
mkdir -p gpurun_out
out=gpurun_out
py=python
t0=$(date +%s)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_stream -s 12 -c 2 -f -o $out/r2w_ncu_k_stream_head \
  $py bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-other-configs > $out/r2w_ncu_head.log 2>&1; echo "ncu k_stream rc=$? [$(( $(date +%s) - t0 )) s]" | tee $out/r2w_summary.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_stream_rows|k_finalize_vbo" -s 300 -c 4 -f -o $out/r2w_ncu_vbo_head \
  $py bench.py --method vb_online --rows 20000000 --batches 10 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-other-configs > $out/r2w_ncu_vbo.log 2>&1; echo "ncu vbo rc=$? [$(( $(date +%s) - t0 )) s]" | tee -a $out/r2w_summary.txt
