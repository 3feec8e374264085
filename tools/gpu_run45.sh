#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
LIB=path_tracer_b200/lib/libptb200.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wave_kernel -s 1 -c 1 -o $O/r2_c4_final -f python tools/variant_time.py $LIB c4 8 1 > $O/r2_ncu_c4_final.log 2>&1
M=smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fp32_pred_on.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:render_wave_kernel -s 1 -c 1 --csv --log-file $O/r2_c4_fp32_counters.csv python tools/variant_time.py $LIB c4 16 1 > /dev/null 2>&1
tail -3 $O/r2_ncu_c4_final.log
