set -u
mkdir -p gpurun_out
timeout 300 python tools/chol_probe.py 2>&1 | grep CHOLPROBE | tee gpurun_out/s3_cholprobe.txt
timeout 600 python -m pytest tests/test_showresults.py tests/test_gpu_bp_batch.py tests/test_gpu_kernels.py tests/test_testers.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/s3_pytest_a.log
timeout 400 python -m pytest tests/test_gpu_baseline_sizes.py -x -q -k "c2" 2>&1 | tail -5 | tee gpurun_out/s3_pytest_c2.log
