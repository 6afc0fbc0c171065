#!/bin/bash
# Round 2, GPU call M (1 GPU): the -m gpu suite at HEAD, vb_online after the cooperative span sums, the hot / cold L1 policy of the
# first-field pass (SVBFM_REC_HOT), then the default bench line (e2e, cpu_baseline, other_configs).
set -u
mkdir -p gpurun_out
out=gpurun_out
py=python
timeout 1500 $py -m pytest tests -m gpu -x -q > $out/r2m_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$? $(tail -1 $out/r2m_pytest_gpu.log)" | tee $out/r2m_summary.txt
run() { name=$1; shift 1;
  timeout 400 env SVBFM_X=0 "$@" $py bench.py $B > $out/r2m_bench_$name.json 2> $out/r2m_bench_$name.err
  echo "bench $name rc=$? $($py - <<PY
import json
try:
    d=json.loads(open("$out/r2m_bench_$name.json").read().strip().splitlines()[-1]); k=d["roofline"]["kernel_classes_ms"]; s=d["steps"]
    print("ms/step %.2f f0 %.2f f1 %.2f fin %.2f clk %s %s" % (d["ms_per_step"], k["stream_v_field0"]/s, k["stream_v_field1"]/s, k["finalize_v"]/s, d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
except Exception as e: print("ERR", e)
PY
)" | tee -a $out/r2m_summary.txt; }
B="--steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-other-configs --method vb_online"
run vbo
B="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs"
run hot0
run hot1024 SVBFM_REC_HOT=1024
run hot2048 SVBFM_REC_HOT=2048
run hot4096 SVBFM_REC_HOT=4096
run hot0_again
timeout 900 $py bench.py --steps 10 --warmup 3 > $out/r2m_bench_default.json 2> $out/r2m_bench_default.err; echo "bench default rc=$? $(grep -o '"ms_per_step": [0-9.]*' $out/r2m_bench_default.json | head -1)" | tee -a $out/r2m_summary.txt
