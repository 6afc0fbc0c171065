#!/bin/bash
# Round 2, GPU call V (1 GPU): k_stream_rows with the head-mask scan that stops at the longest interior column (against r02_r's
# s_rows8: 88.4 ms per 10-batch epoch, passes 16.67 / 17.83), the block cache keyed by handle, the vb_online oracle test at the ML-1M
# shape; then the whole suite, vb_online at 200 M and the default bench line at HEAD.
set -u
mkdir -p gpurun_out
out=gpurun_out
py=python
t0=$(date +%s)
run() { name=$1; flags=$2; shift 2;
  timeout 400 env SVBFM_X=0 "$@" $py bench.py $B $flags > $out/r2v_bench_$name.json 2> $out/r2v_bench_$name.err
  echo "bench $name rc=$? $($py - <<PY
import json
try:
    d=json.loads(open("$out/r2v_bench_$name.json").read().strip().splitlines()[-1]); k=d["roofline"]["kernel_classes_ms"]; s=d["steps"]
    print("ms/step %.3f f0 %.2f f1 %.2f fin %.2f flush %.2f w %.2f launches %d clk %s %s" % (d["ms_per_step"], k["stream_v_field0"]/s, k["stream_v_field1"]/s, k["finalize_v"]/s, k["stream_flush"]/s, (k["stream_w"]+k["finalize_w"])/s, d["gpu_launches"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
except Exception as e: print("ERR", e)
PY
) [$(( $(date +%s) - t0 )) s]" | tee -a $out/r2v_summary.txt; }
: > $out/r2v_summary.txt
B="--steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-other-configs --method vb_online --rows 20000000 --batches 10"
run s_scan ""
timeout 1500 $py -m pytest tests -m gpu -q > $out/r2v_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$? $(tail -1 $out/r2v_pytest_gpu.log) [$(( $(date +%s) - t0 )) s]" | tee -a $out/r2v_summary.txt
B="--steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-other-configs --method vb_online"
run vbo_timed ""
timeout 900 $py bench.py > $out/r2v_bench_default.json 2> $out/r2v_bench_default.err; echo "bench default rc=$? $(grep -o '"ms_per_step": [0-9.]*' $out/r2v_bench_default.json | head -1) [$(( $(date +%s) - t0 )) s]" | tee -a $out/r2v_summary.txt
