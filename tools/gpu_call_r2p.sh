#!/bin/bash
# Round 2, GPU call P (1 GPU): vb_online with the batch in flight packed into contiguous streams (k_vbo_pack) and swept by
# k_stream_rows (default) / by k_stream (SVBFM_VBO_ROWS=0); the vb_online tests on the real GPU; ncu full capture of two batch
# passes of k_stream_rows (a 2 M-entry batch of the same popularity model: --rows 20000000 --batches 10 keeps the launch count low).
set -u
mkdir -p gpurun_out
out=gpurun_out
py=python
timeout 600 $py -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_experiments.py -m gpu -x -q -k "online" > $out/r2p_pytest_vbo.log 2>&1; echo "pytest vb_online rc=$? $(tail -1 $out/r2p_pytest_vbo.log)" | tee $out/r2p_summary.txt
run() { name=$1; shift 1;
  timeout 400 env SVBFM_X=0 "$@" $py bench.py $B > $out/r2p_bench_$name.json 2> $out/r2p_bench_$name.err
  echo "bench $name rc=$? $($py - <<PY
import json
try:
    d=json.loads(open("$out/r2p_bench_$name.json").read().strip().splitlines()[-1]); k=d["roofline"]["kernel_classes_ms"]; s=d["steps"]
    print("ms/step %.2f f0 %.2f f1 %.2f fin %.2f flush %.2f w %.2f clk %s %s" % (d["ms_per_step"], k["stream_v_field0"]/s, k["stream_v_field1"]/s, k["finalize_v"]/s, k["stream_flush"]/s, (k["stream_w"]+k["finalize_w"])/s, d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
except Exception as e: print("ERR", e)
PY
)" | tee -a $out/r2p_summary.txt; }
B="--steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-other-configs --method vb_online"
run vbo_rows
run vbo_packed_kstream SVBFM_VBO_ROWS=0
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_stream_rows|k_finalize_vbo|k_vbo_pack" -s 300 -c 5 -f -o $out/r2p_ncu_vbo_rows \
  $py bench.py --method vb_online --rows 20000000 --batches 10 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-other-configs > $out/r2p_ncu_vbo_rows.log 2>&1; echo "ncu vbo rows rc=$?" | tee -a $out/r2p_summary.txt
