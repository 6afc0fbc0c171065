#!/bin/bash
# Round 2, GPU call R (1 GPU): the -m gpu suite at HEAD (span sums folded into k_finalize, packed vb_online batches with dense column
# ids, means-only rows for the mcmc re-prediction), k_stream_rows occupancy variants on 2 M-entry batches, mcmc at 200 M, the small
# configs, then the default bench line and the ncu launch list of the same command.
set -u
mkdir -p gpurun_out
out=gpurun_out
py=python
t0=$(date +%s)
timeout 1500 $py -m pytest tests -m gpu -x -q > $out/r2r_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$? $(tail -1 $out/r2r_pytest_gpu.log) [$(( $(date +%s) - t0 )) s]" | tee $out/r2r_summary.txt
run() { name=$1; shift 1;
  timeout 400 env SVBFM_X=0 "$@" $py bench.py $B > $out/r2r_bench_$name.json 2> $out/r2r_bench_$name.err
  echo "bench $name rc=$? $($py - <<PY
import json
try:
    d=json.loads(open("$out/r2r_bench_$name.json").read().strip().splitlines()[-1]); k=d["roofline"]["kernel_classes_ms"]; s=d["steps"]
    print("ms/step %.3f f0 %.2f f1 %.2f fin %.2f flush %.2f w %.2f clk %s %s" % (d["ms_per_step"], k["stream_v_field0"]/s, k["stream_v_field1"]/s, k["finalize_v"]/s, k["stream_flush"]/s, (k["stream_w"]+k["finalize_w"])/s, d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
except Exception as e: print("ERR", e)
PY
) [$(( $(date +%s) - t0 )) s]" | tee -a $out/r2r_summary.txt; }
B="--steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-other-configs --method vb_online --rows 20000000 --batches 10"
run s_rows8
run s_rows10 SVBFM_LIB=build/variants/libsvbfm_rows10.so
run s_rows12 SVBFM_LIB=build/variants/libsvbfm_rows12.so
B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-other-configs --workload ml1m"
run ml1m_fold
run ml1m_nofold SVBFM_SPAN_FOLD=0
B="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs --workload ml10m"
run ml10m_fold
run ml10m_nofold SVBFM_SPAN_FOLD=0
B="--steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs --method mcmc"
run mcmc_200m
B="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs"
run vb_nofold SVBFM_SPAN_FOLD=0
timeout 900 $py bench.py --steps 10 --warmup 3 > $out/r2r_bench_default.json 2> $out/r2r_bench_default.err; echo "bench default rc=$? $(grep -o '"ms_per_step": [0-9.]*' $out/r2r_bench_default.json | head -1) [$(( $(date +%s) - t0 )) s]" | tee -a $out/r2r_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $out/r2r_launches_kdd200m_k50.csv \
  $py bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-other-configs > $out/r2r_ncu_launches.log 2>&1; echo "ncu launch list rc=$? [$(( $(date +%s) - t0 )) s]" | tee -a $out/r2r_summary.txt
