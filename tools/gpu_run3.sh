#!/bin/bash
# round-2 GPU session 3: C1 after the kTrees split; first bench lines of c2, c3, c5 (short)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for v in r1 path_tracer_b200/lib/libptb200; do
  f=$v.so; [ -f build/variants/$v.so ] && f=build/variants/$v.so
  timeout 120 python tools/variant_time.py $f c1 100 5 >> gpurun_out/r2_variants3.log 2>&1
done
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest3.log 2>&1
timeout 300 python bench.py --workload c2 --steps 3 --warmup 3 > gpurun_out/r2_bench_c2_a.log 2>&1
timeout 300 python bench.py --workload c3 --steps 2 --warmup 3 > gpurun_out/r2_bench_c3_a.log 2>&1
timeout 600 python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c5_a.log 2>&1
cat gpurun_out/r2_variants3.log; tail -2 gpurun_out/r2_pytest3.log
timeout 400 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c4_b.log 2>&1
tail -c 300 gpurun_out/r2_bench_c4_b.log
for c in c2 c3 c5; do tail -c 400 gpurun_out/r2_bench_${c}_a.log; echo; done
