#!/bin/bash
# One GPU visit: tests, bench (both arms), ncu launch list + full capture of the two dominant kernels.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --light"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f64_dmma -s 0 -c 1 -o gpurun_out/prof_gram -f $CMD > gpurun_out/ncu_gram.log 2>&1
echo "ncu gram rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:coldot -s 41 -c 2 -o gpurun_out/prof_coldot -f $CMD > gpurun_out/ncu_coldot.log 2>&1
echo "ncu coldot rc=$?"
ls -la gpurun_out
