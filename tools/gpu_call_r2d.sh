#!/bin/bash
# Round 2, GPU call D: k_stream's TMA loop rewritten for instruction count (shared addresses formed once, slots of two batches,
# carried slot / parity): ring tests, bench, slot-size / ring-depth variants, ncu full capture.
set -u
mkdir -p gpurun_out
out=gpurun_out
py=python
B="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs"
timeout 600 $py -m pytest tests/test_gpu_zzz_tma_ring.py tests/test_gpu_parity.py -m gpu -x -q > $out/r2d_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 $out/r2d_pytest.log)" | tee $out/r2d_summary.txt
run() { name=$1; lib=$2; shift 2; L=""; [ "$lib" != base ] && L="SVBFM_LIB=build/variants/libsvbfm_$lib.so";
  timeout 300 env SVBFM_X=0 $L "$@" $py bench.py $B > $out/r2d_bench_$name.json 2> $out/r2d_bench_$name.err
  echo "bench $name rc=$? $($py - <<PY
import json
try:
    d=json.loads(open("$out/r2d_bench_$name.json").read().strip().splitlines()[-1]); k=d["roofline"]["kernel_classes_ms"]; s=d["steps"]
    print("ms/step %.2f f0 %.2f f1 %.2f fin %.2f clk %s %s" % (d["ms_per_step"], k["stream_v_field0"]/s, k["stream_v_field1"]/s, k["finalize_v"]/s, d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
except Exception as e: print("ERR", e)
PY
)" | tee -a $out/r2d_summary.txt; }
run base base
run nst2 nst2
run nst4 nst4
run sb1nst4 sb1nst4
run sb4nst2 sb4nst2
run sb4nst3 sb4nst3
run base_again base
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stream -s 12 -c 2 -f -o $out/r2d_ncu_k_stream \
  $py bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-other-configs > $out/r2d_ncu.log 2>&1; echo "ncu full rc=$?" | tee -a $out/r2d_summary.txt
