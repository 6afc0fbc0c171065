#!/bin/bash
# Round 2, GPU call C: the whole -m gpu suite at HEAD (rank layout + TMA ring + 4x7 CTAs are the defaults now), the full bench line,
# ring-depth / tile-size variants of k_stream, the ncu launch list and a full capture of k_stream (both fields) at HEAD.
set -u
mkdir -p gpurun_out
out=gpurun_out
py=python
B="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 1200 $py -m pytest tests -m gpu -x -q > $out/r2c_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$? $(tail -1 $out/r2c_pytest_gpu.log)" | tee $out/r2c_summary.txt
timeout 600 $py bench.py --steps 10 --warmup 3 > $out/r2c_bench_head.json 2> $out/r2c_bench_head.err; echo "bench head rc=$? $(grep -o '"ms_per_step": [0-9.]*' $out/r2c_bench_head.json | head -1)" | tee -a $out/r2c_summary.txt
run() { name=$1; lib=$2; shift 2; L=""; [ "$lib" != base ] && L="SVBFM_LIB=build/variants/libsvbfm_$lib.so";
  timeout 300 env SVBFM_X=0 $L "$@" $py bench.py $B > $out/r2c_bench_$name.json 2> $out/r2c_bench_$name.err
  echo "bench $name rc=$? $($py - <<PY
import json
try:
    d=json.loads(open("$out/r2c_bench_$name.json").read().strip().splitlines()[-1]); k=d["roofline"]["kernel_classes_ms"]; s=d["steps"]
    print("ms/step %.2f f0 %.2f f1 %.2f fin %.2f clk %s %s" % (d["ms_per_step"], k["stream_v_field0"]/s, k["stream_v_field1"]/s, k["finalize_v"]/s, d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
except Exception as e: print("ERR", e)
PY
)" | tee -a $out/r2c_summary.txt; }
run base base
run nst4 nst4
run nst3 nst3
run nst8 nst8
run tile2048 base SVBFM_TILE_ENTRIES=2048
run tile8192 base SVBFM_TILE_ENTRIES=8192
run w2m14 w2m14
run w8m3 w8m3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/r2c_launches.csv $py bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > $out/r2c_ncu_launches.log 2>&1; echo "ncu launches rc=$?" | tee -a $out/r2c_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stream -s 12 -c 2 -f -o $out/r2c_ncu_k_stream_head \
  $py bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $out/r2c_ncu_head.log 2>&1; echo "ncu full rc=$?" | tee -a $out/r2c_summary.txt
