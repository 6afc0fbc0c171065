#!/bin/bash
# Round 2, GPU call Q (1 GPU): vb_online with dense column ids + carried finalize operands + batch predictions over the transposed
# parameters (all default now), each switched off in turn; k_stream_rows tile sizes and the gather-one-row-ahead variant on a
# 20 M-rating set with the same 2 M-entry batches (--rows 20000000 --batches 10).
set -u
mkdir -p gpurun_out
out=gpurun_out
py=python
timeout 600 $py -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_experiments.py -m gpu -x -q -k "online" > $out/r2q_pytest_vbo.log 2>&1; echo "pytest vb_online rc=$? $(tail -1 $out/r2q_pytest_vbo.log)" | tee $out/r2q_summary.txt
run() { name=$1; shift 1;
  timeout 400 env SVBFM_X=0 "$@" $py bench.py $B > $out/r2q_bench_$name.json 2> $out/r2q_bench_$name.err
  echo "bench $name rc=$? $($py - <<PY
import json
try:
    d=json.loads(open("$out/r2q_bench_$name.json").read().strip().splitlines()[-1]); k=d["roofline"]["kernel_classes_ms"]; s=d["steps"]
    print("ms/step %.2f f0 %.2f f1 %.2f fin %.2f flush %.2f w %.2f clk %s %s" % (d["ms_per_step"], k["stream_v_field0"]/s, k["stream_v_field1"]/s, k["finalize_v"]/s, k["stream_flush"]/s, (k["stream_w"]+k["finalize_w"])/s, d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
except Exception as e: print("ERR", e)
PY
)" | tee -a $out/r2q_summary.txt; }
B="--steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-other-configs --method vb_online"
run vbo_default
run vbo_global_ids SVBFM_VBO_COMPACT=0
run vbo_predict_kd SVBFM_VBO_PREDICT2=0
B="--steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-other-configs --method vb_online --rows 20000000 --batches 10"
run s_default
run s_tile128 SVBFM_TILE_ENTRIES=128
run s_tile512 SVBFM_TILE_ENTRIES=512
run s_tile1024 SVBFM_TILE_ENTRIES=1024
run s_pipe SVBFM_LIB=build/variants/libsvbfm_rpipe.so
run s_pipe_tile512 SVBFM_LIB=build/variants/libsvbfm_rpipe.so SVBFM_TILE_ENTRIES=512
