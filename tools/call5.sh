set -u
mkdir -p gpurun_out
timeout 600 python tools/chol_probe.py 2>&1 | grep CHOLPROBE | tee gpurun_out/s6_cholprobe.txt
