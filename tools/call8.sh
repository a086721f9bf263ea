set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/s2_topo.txt 2>&1
timeout 400 $TR --nproc-per-node 8 --master-port 29621 bench.py --gpus 8 --steps 5 2>gpurun_out/s2_bench8.err | grep '^{' | tee gpurun_out/s2_bench_n8.json | cut -c1-300
timeout 400 $TR --nproc-per-node 4 --master-port 29622 bench.py --gpus 4 --steps 5 2>gpurun_out/s2_bench4.err | grep '^{' | tee gpurun_out/s2_bench_n4.json | cut -c1-300
ADMM_B200_PERSIST_PROF=1 timeout 300 $TR --nproc-per-node 8 --master-port 29623 bench.py --gpus 8 --only-svm > gpurun_out/s2_svm_prof_n8.txt 2>&1
grep svm_c3 gpurun_out/s2_svm_prof_n8.txt | cut -c1-1500
tail -3 gpurun_out/s2_bench8.err gpurun_out/s2_bench4.err
