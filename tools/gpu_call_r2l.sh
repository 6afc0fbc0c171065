#!/bin/bash
# Round 2, GPU call L (1 GPU): record gathers of the second-field pass past L1 (the default now), vb_online with the finalize walking a
# batch's non-empty columns only, MCMC.
set -u
mkdir -p gpurun_out
out=gpurun_out
py=python
B="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs"
run() { name=$1; lib=$2; shift 2; L=""; [ "$lib" != base ] && L="SVBFM_LIB=build/variants/libsvbfm_$lib.so";
  timeout 400 env SVBFM_X=0 $L "$@" $py bench.py $B > $out/r2l_bench_$name.json 2> $out/r2l_bench_$name.err
  echo "bench $name rc=$? $($py - <<PY
import json
try:
    d=json.loads(open("$out/r2l_bench_$name.json").read().strip().splitlines()[-1]); k=d["roofline"]["kernel_classes_ms"]; s=d["steps"]
    print("ms/step %.2f f0 %.2f f1 %.2f fin %.2f clk %s %s" % (d["ms_per_step"], k["stream_v_field0"]/s, k["stream_v_field1"]/s, k["finalize_v"]/s, d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
except Exception as e: print("ERR", e)
PY
)" | tee -a $out/r2l_summary.txt; }
: > $out/r2l_summary.txt
run na2 base
run na0 base SVBFM_REC_NA=0
run nst4_na2 nst4
run na2_again base
B="--steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-other-configs --method vb_online"
run vbo_clist base
run vbo_noclist base SVBFM_VBO_NO_CLIST=1
B="--steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-other-configs --method mcmc"
run mcmc base
B="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs --workload netflix"
run netflix base
