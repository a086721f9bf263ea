#!/bin/bash
# One single-GPU visit (round 2): the whole GPU test suite, smoke(), both bench arms, the per-config measurements, then the
# profile refresh of tools/profile_round.sh (launch list + ncu --set full of the dominant kernels).  Output: gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
timeout 300 python tools/bench_configs.py 2>&1 | grep CONFIGS | tee gpurun_out/configs.log
for p in huber lad; do timeout 300 python tools/bench_unwrapped.py --problem $p --iters 100 2>&1 | grep UNWRAPPED; done | tee gpurun_out/unwrapped_1gpu.log
timeout 300 python tools/chol_probe.py 2>&1 | grep CHOLPROBE | tee gpurun_out/cholprobe.txt
bash tools/profile_round.sh 2>&1 | tail -30
