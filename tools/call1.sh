set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_lasso.py tests/test_gpu_conditioning.py tests/test_gpu_model.py tests/test_gpu_bp_batch.py -x -q 2>&1 | tail -5 | tee gpurun_out/s2_pytest_inv.log
timeout 300 python -m pytest tests/test_gpu_baseline_sizes.py -x -q -k "c2" 2>&1 | tail -5 | tee gpurun_out/s2_pytest_c2.log
timeout 300 python bench.py --steps 3 --no-cpu --no-svm 2>gpurun_out/s2_b1.err | tee gpurun_out/s2_bench_n1_inv.json
ADMM_B200_NO_INV_OVERLAP=1 timeout 300 python bench.py --steps 3 --no-cpu --no-svm --no-e2e --light 2>gpurun_out/s2_b1b.err | tee gpurun_out/s2_bench_n1_noinv.json
