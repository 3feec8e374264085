#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
{
python tools/variant_time.py path_tracer_b200/lib/libptb200.so c4 16 3
python tools/variant_time.py path_tracer_b200/lib/libptb200.so c1 100 4
timeout 900 python -m pytest tests -m gpu -x -q -k "flat or config4 or mesh or tree" 2>&1 | tail -3
timeout 600 python tools/nonfinite_pixels.py c5
timeout 600 python tools/nonfinite_pixels.py c3
timeout 600 python tools/nonfinite_pixels.py c4 32
} > $O/r2_run21.log 2>&1
M=smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fp32_pred_on.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
for c in "c2 64" "c3 64" "c5 16" "c4 16"; do set -- $c
timeout 600 ncu --metrics $M --clock-control none -k regex:render_wave_kernel -s 1 -c 1 --csv --log-file $O/r2_$1_fp32_counters.csv python tools/variant_time.py path_tracer_b200/lib/libptb200.so $1 $2 1 > /dev/null 2>&1
done
cat $O/r2_run21.log
