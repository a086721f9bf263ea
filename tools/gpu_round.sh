#!/bin/bash
# One GPU visit: tests, bench (both arms), ncu launch list + full captures of the dominant kernels,
# and the per-config measurements (C1, C3, C4, C5a, C5b).  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
timeout 300 python tools/bench_configs.py 2>&1 | grep CONFIGS | tee gpurun_out/configs.log
for p in svm huber lad; do timeout 300 python tools/bench_unwrapped.py --problem $p --iters 100 2>&1 | grep UNWRAPPED; done | tee gpurun_out/unwrapped_1gpu.log
timeout 200 python tools/bench_graph.py 2>&1 | grep GRAPH | tee gpurun_out/bench_graph.log
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --light"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f64_dmma -s 0 -c 1 -o gpurun_out/prof_gram -f $CMD > gpurun_out/ncu_gram.log 2>&1
echo "ncu gram rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:coldot -s 41 -c 2 -o gpurun_out/prof_coldot -f $CMD > gpurun_out/ncu_coldot.log 2>&1
echo "ncu coldot rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tv_fused -s 20 -c 1 -o gpurun_out/prof_tvfused -f python tools/bench_configs.py c5a > gpurun_out/ncu_tvfused.log 2>&1
echo "ncu tvfused rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:uw_onepass_kernel -s 20 -c 1 -o gpurun_out/prof_onepass -f python tools/bench_unwrapped.py --problem svm --iters 20 > gpurun_out/ncu_onepass.log 2>&1
echo "ncu onepass rc=$?"
ls -la gpurun_out | head -50
