set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_lasso.py -x -q 2>&1 | tail -3 | tee gpurun_out/s11_pytest.log
for c in 1024 2048 4096 16384; do
  echo "chunk $c"; ADMM_B200_SYMTRI_CHUNK=$c timeout 200 python bench.py --steps 2 --no-cpu --no-svm --no-e2e --light 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['loop_us_per_iter'])"
done 2>&1 | tee gpurun_out/s11_chunks.txt
