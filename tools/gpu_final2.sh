#!/bin/bash
# round-2 final evidence on one GPU: the whole GPU test suite, bench lines of every workload, the reference arm,
# ncu launch list + full captures + instruction counters, compute-sanitizer
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out; O=gpurun_out
LIB=path_tracer_b200/lib/libptb200.so
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $O/r2_final_pytest_gpu.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2_final_smoke.log 2>&1
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > $O/r2_final_c1.json 2> $O/r2_final_c1.err
timeout 900 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > $O/r2_final_ref_c1.json 2> $O/r2_final_ref_c1.err
for c in c2 c3 c4 c5; do timeout 900 python bench.py --workload $c > $O/r2_final_$c.json 2> $O/r2_final_$c.err; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2_bench_under_ncu.log 2>&1
M=smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fp32_pred_on.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
for c in "c1 100" "c2 64" "c3 64" "c4 16" "c5 16"; do set -- $c
timeout 600 ncu --metrics $M --clock-control none -k regex:render_wave_kernel -s 1 -c 1 --csv --log-file $O/r2_$1_fp32_counters.csv python tools/variant_time.py $LIB $1 $2 1 > /dev/null 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wave_kernel -s 1 -c 1 -o $O/r2_c1_final -f python tools/variant_time.py $LIB c1 100 1 > $O/r2_ncu_c1_final.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_wave_kernel -s 1 -c 1 -o $O/r2_c4_final -f python tools/variant_time.py $LIB c4 8 1 > $O/r2_ncu_c4_final.log 2>&1
{
echo "--- racecheck (wavefront kernel)"; PT_TINY_LPT=1 timeout 1200 compute-sanitizer --tool racecheck python tools/tiny_render.py 2>&1 | grep -v "^=========\s*$" | tail -14
echo "--- memcheck (wavefront kernel)"; PT_TINY_LPT=1 timeout 1200 compute-sanitizer --tool memcheck python tools/tiny_render.py 2>&1 | tail -4
echo "--- synccheck (wavefront kernel)"; timeout 1200 compute-sanitizer --tool synccheck python tools/tiny_render.py 2>&1 | tail -2
echo "--- memcheck (lane kernel)"; timeout 1200 compute-sanitizer --tool memcheck python tools/tiny_render.py 1 2>&1 | tail -2
} > $O/r2_final_sanitizer.log 2>&1
tail -4 $O/r2_final_pytest_gpu.log; cat $O/r2_final_smoke.log | tail -2
for f in c1 ref_c1 c2 c3 c4 c5; do echo "== $f"; head -c 300 $O/r2_final_$f.json; echo; tail -n 2 $O/r2_final_$f.err; done
tail -20 $O/r2_final_sanitizer.log
