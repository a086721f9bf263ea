set -u
mkdir -p gpurun_out
ADMM_B200_SYMTRI_PROBE=1 timeout 400 python bench.py --steps 2 --no-cpu --no-svm --no-e2e --light 2>gpurun_out/s12_b1.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('probe', d['loop_us_per_iter'])"
timeout 400 python bench.py --steps 2 --no-cpu --no-svm --no-e2e --light 2>gpurun_out/s12_b1.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('real ', d['loop_us_per_iter'])"
