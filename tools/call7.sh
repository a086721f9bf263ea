set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fast.py tests/test_gpu_tv.py tests/test_gpu_graph.py -x -q 2>&1 | tail -12 | tee gpurun_out/s8_pytest_tv.log
