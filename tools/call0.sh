set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 500 python -m pytest tests/test_gpu_sharded.py -x -q -k "lasso or persist" 2>&1 | tail -5 | tee gpurun_out/s2_pytest_2gpu.log
timeout 300 python bench.py --only-svm 2>gpurun_out/s2_svm1.err | tee gpurun_out/s2_svm_n1.json
timeout 300 $TR --nproc-per-node 2 --master-port 29611 bench.py --gpus 2 --only-svm 2>gpurun_out/s2_svm2.err | tee gpurun_out/s2_svm_n2.json
timeout 400 $TR --nproc-per-node 2 --master-port 29612 bench.py --gpus 2 --steps 3 --no-cpu 2>gpurun_out/s2_bench2.err | tee gpurun_out/s2_bench_n2.json
tail -3 gpurun_out/s2_bench2.err
