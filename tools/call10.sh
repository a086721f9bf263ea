set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_lasso.py tests/test_gpu_model.py tests/test_gpu_unwrapped.py -x -q 2>&1 | tail -4 | tee gpurun_out/s10_pytest.log
timeout 400 python bench.py --steps 3 --no-cpu --no-svm --no-e2e 2>gpurun_out/s10_b1.err | tee gpurun_out/s10_bench_n1.json | cut -c1-100
ADMM_B200_NO_SYMTRI=1 timeout 200 python tools/bench_configs.py c1 2>&1 | grep CONFIGS | cut -c1-400 | tee gpurun_out/s10_c1_nosymtri.log
timeout 200 python tools/bench_configs.py c1 2>&1 | grep CONFIGS | cut -c1-400 | tee gpurun_out/s10_c1.log
