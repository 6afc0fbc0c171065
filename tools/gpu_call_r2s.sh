#!/bin/bash
# Round 2, GPU call S (1 GPU): the small configs with and without the per-class event pairs inside the timed region, tile sizes at the
# ML-10M shape, the vb_online tests again (tolerance of the dense-id comparison), vb_online at HEAD.
set -u
mkdir -p gpurun_out
out=gpurun_out
py=python
t0=$(date +%s)
timeout 600 $py -m pytest tests/test_gpu_zz_experiments.py -m gpu -x -q > $out/r2s_pytest_zz.log 2>&1; echo "pytest zz_experiments rc=$? $(tail -1 $out/r2s_pytest_zz.log) [$(( $(date +%s) - t0 )) s]" | tee $out/r2s_summary.txt
run() { name=$1; flags=$2; shift 2;
  timeout 400 env SVBFM_X=0 "$@" $py bench.py $B $flags > $out/r2s_bench_$name.json 2> $out/r2s_bench_$name.err
  echo "bench $name rc=$? $($py - <<PY
import json
try:
    d=json.loads(open("$out/r2s_bench_$name.json").read().strip().splitlines()[-1]); k=d["roofline"]["kernel_classes_ms"]; s=d["steps"]
    print("ms/step %.3f f0 %.2f f1 %.2f fin %.2f flush %.2f w %.2f launches %d clk %s %s" % (d["ms_per_step"], k["stream_v_field0"]/s, k["stream_v_field1"]/s, k["finalize_v"]/s, k["stream_flush"]/s, (k["stream_w"]+k["finalize_w"])/s, d["gpu_launches"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
except Exception as e: print("ERR", e)
PY
) [$(( $(date +%s) - t0 )) s]" | tee -a $out/r2s_summary.txt; }
B="--steps 400 --warmup 20 --no-cpu-baseline --no-e2e --no-other-configs --workload ml1m"
run ml1m_timed ""
run ml1m_after "--profile after"
run ml1m_off "--profile off"
run ml1m_tile128 "--profile after" SVBFM_TILE_ENTRIES=128
run ml1m_tile512 "--profile after" SVBFM_TILE_ENTRIES=512
B="--steps 60 --warmup 5 --no-cpu-baseline --no-e2e --no-other-configs --workload ml10m"
run ml10m_timed ""
run ml10m_after "--profile after"
run ml10m_tile1024 "--profile after" SVBFM_TILE_ENTRIES=1024
run ml10m_tile2048 "--profile after" SVBFM_TILE_ENTRIES=2048
run ml10m_tile256 "--profile after" SVBFM_TILE_ENTRIES=256
B="--steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-other-configs --method vb_online"
run vbo_after "--profile after"
