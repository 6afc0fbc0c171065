#!/bin/bash
# Round 2, GPU call K (1 GPU): the -m gpu suite at HEAD (x = NULL, CLI cross shards, padded fallback), then the cache policy of k_stream's record gather:
# plain LDG.256 (base), ld.global.nc (rec1), nc + L1::evict_last (rec2), nc + L1::no_allocate (rec3).
set -u
mkdir -p gpurun_out
out=gpurun_out
py=python
B="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs"
timeout 1500 $py -m pytest tests -m gpu -x -q > $out/r2k_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$? $(tail -1 $out/r2k_pytest_gpu.log)" | tee $out/r2k_summary.txt
run() { name=$1; lib=$2; shift 2; L=""; [ "$lib" != base ] && L="SVBFM_LIB=build/variants/libsvbfm_$lib.so";
  timeout 300 env SVBFM_X=0 $L "$@" $py bench.py $B > $out/r2k_bench_$name.json 2> $out/r2k_bench_$name.err
  echo "bench $name rc=$? $($py - <<PY
import json
try:
    d=json.loads(open("$out/r2k_bench_$name.json").read().strip().splitlines()[-1]); k=d["roofline"]["kernel_classes_ms"]; s=d["steps"]
    print("ms/step %.2f f0 %.2f f1 %.2f fin %.2f clk %s %s" % (d["ms_per_step"], k["stream_v_field0"]/s, k["stream_v_field1"]/s, k["finalize_v"]/s, d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
except Exception as e: print("ERR", e)
PY
)" | tee -a $out/r2k_summary.txt; }
run base base
run rec1 rec1
run rec2 rec2
run rec3 rec3
run base_again base
timeout 300 $py bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-other-configs > $out/r2k_bench_e2e.json 2> $out/r2k_bench_e2e.err; echo "bench e2e (x = NULL) rc=$? $($py -c "import json;d=json.loads(open('$out/r2k_bench_e2e.json').read().strip().splitlines()[-1]);print(d['e2e']['ms_per_step'], d['e2e']['last_step_phases_ms'], d['e2e']['h2d_bytes_per_step'])")" | tee -a $out/r2k_summary.txt
