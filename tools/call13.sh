set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_lasso.py tests/test_gpu_unwrapped.py -x -q 2>&1 | tail -3 | tee gpurun_out/s13_pytest.log
timeout 300 python tools/bench_batch.py 64 2>&1 | grep BATCH | tee gpurun_out/r02b_batch_plain.log
timeout 600 ncu --set full --clock-control none -k regex:gemm_f64_dmma_kernel -s 340 -c 2 -o gpurun_out/r02b_lambda_batch -f python tools/bench_batch.py 64 > gpurun_out/r02b_ncu_batch.log 2>&1
echo "ncu lambda batch rc=$?"
python tools/extract_reports.py r02c "python tools/bench_batch.py 64 (launches 340-341 of gemm_f64_dmma_kernel: the two triangular products of one batch iteration)" -- gpurun_out/r02b_lambda_batch.ncu-rep > gpurun_out/r02c_ncu_summary.txt 2>&1
rm -f gpurun_out/r02b_lambda_batch.ncu-rep
grep -E "Kernel Name|Grid Size|time_duration|dmma_cycles|dram__bytes_read" gpurun_out/r02c_ncu_summary.txt | cut -c1-200
