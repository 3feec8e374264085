#!/bin/bash
# evidence of the build with demand-following item-list shares: GPU tests, smoke, bench lines of all workloads,
# instruction counters of c2 / c5 / c1 (traffic.json), full capture of c2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
LIB=path_tracer_b200/lib/libptb200.so
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2b_pytest_gpu.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2b_smoke.log 2>&1
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > $O/r2b_c1.json 2> $O/r2b_c1.err
for c in c2 c3 c4 c5; do timeout 900 python bench.py --workload $c > $O/r2b_$c.json 2> $O/r2b_$c.err; done
M=smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fp32_pred_on.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
for c in "c1 100" "c2 64" "c5 16"; do set -- $c
timeout 600 ncu --metrics $M --clock-control none -k regex:render_wave_kernel -s 1 -c 1 --csv --log-file $O/r2b_$1_fp32_counters.csv python tools/variant_time.py $LIB $1 $2 1 > /dev/null 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wave_kernel -s 1 -c 1 -o $O/r2b_c2_final -f python tools/variant_time.py $LIB c2 16 1 > $O/r2b_ncu_c2_final.log 2>&1
tail -3 $O/r2b_pytest_gpu.log; tail -2 $O/r2b_smoke.log
for f in c1 c2 c3 c4 c5; do echo "== $f"; head -c 260 $O/r2b_$f.json; echo; tail -n 2 $O/r2b_$f.err; done
