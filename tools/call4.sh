set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/s5_pytest_gpu.log
timeout 300 python tools/bench_configs.py 2>&1 | grep CONFIGS | tee gpurun_out/s5_configs.log
for p in huber lad; do timeout 300 python tools/bench_unwrapped.py --problem $p --iters 100 2>&1 | grep UNWRAPPED; done | tee gpurun_out/s5_unwrapped_1gpu.log
timeout 400 python bench.py --steps 3 --no-cpu --no-svm 2>gpurun_out/s5_b1.err | tee gpurun_out/s5_bench_n1.json | cut -c1-120
