set -u
mkdir -p gpurun_out
timeout 600 python tools/chol_probe.py 2>&1 | grep CHOLPROBE | tee gpurun_out/s6_cholprobe.txt
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_lasso.py tests/test_gpu_model.py -x -q 2>&1 | tail -4 | tee gpurun_out/s6_pytest.log
