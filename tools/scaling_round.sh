# One 8-GPU box: bench.py at N = 8, 4, 2, 1 back to back (the curve of profiles/r02_scaling.txt) and the two-rank parity cases.
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 8 --master-port 29621 bench.py --gpus 8 --steps 5 2>gpurun_out/s9_bench8.err | grep '^{' | tee gpurun_out/s9_bench_n8.json | cut -c1-200
timeout 400 $TR --nproc-per-node 4 --master-port 29622 bench.py --gpus 4 --steps 5 2>gpurun_out/s9_bench4.err | grep '^{' | tee gpurun_out/s9_bench_n4.json | cut -c1-200
timeout 400 $TR --nproc-per-node 2 --master-port 29623 bench.py --gpus 2 --steps 5 2>gpurun_out/s9_bench2.err | grep '^{' | tee gpurun_out/s9_bench_n2.json | cut -c1-200
timeout 400 python bench.py --steps 5 --no-cpu 2>gpurun_out/s9_bench1.err | grep '^{' | tee gpurun_out/s9_bench_n1.json | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -4 | tee gpurun_out/s9_pytest_sharded.log
for f in gpurun_out/s9_bench8.err gpurun_out/s9_bench4.err gpurun_out/s9_bench2.err gpurun_out/s9_bench1.err; do tail -n 2 $f; done
