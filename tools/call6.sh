set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_conditioning.py -x -q 2>&1 | tail -4 | tee gpurun_out/s7_pytest.log
timeout 300 python -m pytest tests/test_gpu_baseline_sizes.py -x -q -k "c2_factor" 2>&1 | tail -4 | tee gpurun_out/s7_pytest_c2.log
timeout 400 python bench.py --steps 3 --no-cpu --no-svm --no-e2e 2>gpurun_out/s7_b1.err | tee gpurun_out/s7_bench_n1.json | cut -c1-100
tail -3 gpurun_out/s7_b1.err
