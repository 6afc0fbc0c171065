#!/bin/bash
# Round 2, GPU call X (1 GPU): the cache / staging choices of k_stream that were tuned at 200 M, at the ML-1M and ML-10M shapes
# (bulk-copy ring against plain loads, L1 policy of the record gathers), timed steps uninstrumented.
set -u
mkdir -p gpurun_out
out=gpurun_out
py=python
t0=$(date +%s)
: > $out/r2x_summary.txt
run() { name=$1; shift 1;
  timeout 200 env SVBFM_X=0 "$@" $py bench.py $B --profile after > $out/r2x_bench_$name.json 2> $out/r2x_bench_$name.err
  echo "bench $name rc=$? $($py - <<PY
import json
try:
    d=json.loads(open("$out/r2x_bench_$name.json").read().strip().splitlines()[-1]); k=d["roofline"]["kernel_classes_ms"]; s=d["steps"]
    print("ms/step %.3f f0 %.2f f1 %.2f fin %.2f clk %s %s" % (d["ms_per_step"], k["stream_v_field0"]/s, k["stream_v_field1"]/s, k["finalize_v"]/s, d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
except Exception as e: print("ERR", e)
PY
) [$(( $(date +%s) - t0 )) s]" | tee -a $out/r2x_summary.txt; }
B="--steps 400 --warmup 20 --no-cpu-baseline --no-e2e --no-other-configs --workload ml1m"
run ml1m_default
run ml1m_plain_loads SVBFM_STREAM_TMA=0
run ml1m_rec_na0 SVBFM_REC_NA=0
run ml1m_rec_hot0 SVBFM_REC_HOT=0
B="--steps 60 --warmup 5 --no-cpu-baseline --no-e2e --no-other-configs --workload ml10m"
run ml10m_default
run ml10m_plain_loads SVBFM_STREAM_TMA=0
run ml10m_rec_na0 SVBFM_REC_NA=0
run ml10m_rec_hot0 SVBFM_REC_HOT=0
