#!/bin/bash
# Round 2, GPU call B: k_stream tuning variants (records gathered one batch ahead over the TMA ring, CTA size, ring depth, prefetch)
set -u
mkdir -p gpurun_out
out=gpurun_out
py=python
B="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
SVBFM_RUN_EXPERIMENTS=1 timeout 300 $py -m pytest tests/test_gpu_zzz_tma_ring.py -m gpu -q -x > $out/r2b_pytest_tma_gp.log 2>&1; echo "tma+gp ring tests rc=$?" | tee $out/r2b_summary.txt
run() { name=$1; lib=$2; shift 2; L=""; [ "$lib" != base ] && L="SVBFM_LIB=build/variants/libsvbfm_$lib.so";
  timeout 300 env SVBFM_REC_RANK=1 $L "$@" $py bench.py $B > $out/r2b_bench_$name.json 2> $out/r2b_bench_$name.err
  echo "bench $name rc=$? $($py - <<PY
import json
try:
    d=json.loads(open("$out/r2b_bench_$name.json").read().strip().splitlines()[-1]); k=d["roofline"]["kernel_classes_ms"]; s=d["steps"]
    print("ms/step %.2f f0 %.2f f1 %.2f fin %.2f clk %s %s" % (d["ms_per_step"], k["stream_v_field0"]/s, k["stream_v_field1"]/s, k["finalize_v"]/s, d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
except Exception as e: print("ERR", e)
PY
)" | tee -a $out/r2b_summary.txt; }
run base_tma base SVBFM_STREAM_TMA=1
run base_tma_gp base SVBFM_STREAM_TMA=1 SVBFM_STREAM_GP=1
run w4m6_tma_gp w4m6 SVBFM_STREAM_TMA=1 SVBFM_STREAM_GP=1
run w4m7_tma w4m7 SVBFM_STREAM_TMA=1
run w4m7_plain w4m7 SVBFM_STREAM_TMA=0
run w4m7_tma_gp w4m7 SVBFM_STREAM_TMA=1 SVBFM_STREAM_GP=1
run nst8w4_tma_gp nst8w4 SVBFM_STREAM_TMA=1 SVBFM_STREAM_GP=1
run nst4_tma_gp nst4 SVBFM_STREAM_TMA=1 SVBFM_STREAM_GP=1
run pf1_tma pf1 SVBFM_STREAM_TMA=1
run pf2_tma pf2 SVBFM_STREAM_TMA=1
run u1m4_tma_gp u1m4 SVBFM_STREAM_TMA=1 SVBFM_STREAM_GP=1
run u1m4_tma u1m4 SVBFM_STREAM_TMA=1
run u1m5_tma_gp u1m5 SVBFM_STREAM_TMA=1 SVBFM_STREAM_GP=1
