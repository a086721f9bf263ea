set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lasso.py -x -q 2>&1 | tail -5 | tee gpurun_out/s4_pytest_lasso.log
timeout 400 python bench.py --steps 3 --no-cpu --no-svm 2>gpurun_out/s4_b1.err | tee gpurun_out/s4_bench_n1.json | cut -c1-200
tail -3 gpurun_out/s4_b1.err
