# Profile refresh (ONE GPU): launch list of the bench command + ncu --set full of the dominant kernels.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --light"
timeout 300 $CMD > gpurun_out/r02b_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r02b_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02b_launches.csv $CMD > gpurun_out/r02b_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tma_kernel -s 3 -c 1 -o gpurun_out/r02b_gram -f $CMD > gpurun_out/r02b_ncu_gram.log 2>&1
echo "ncu gram rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:symtri_kernel -s 60 -c 1 -o gpurun_out/r02b_symtri -f $CMD > gpurun_out/r02b_ncu_symtri.log 2>&1
echo "ncu symtri rc=$?"
timeout 300 python tools/bench_batch.py 64 > gpurun_out/r02b_batch_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_f64_dmma_kernel.*256 -s 6 -c 2 -o gpurun_out/r02b_lambda_batch -f python tools/bench_batch.py 64 > gpurun_out/r02b_ncu_batch.log 2>&1
echo "ncu lambda batch rc=$?"
timeout 300 python bench.py --only-svm > gpurun_out/r02b_svm_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:uw_persist_kernel -s 1 -c 1 -o gpurun_out/r02b_persist -f python bench.py --only-svm > gpurun_out/r02b_ncu_persist.log 2>&1
echo "ncu persist rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:uwb_persist_kernel -s 1 -c 1 -o gpurun_out/r02b_persist_batch -f python bench.py --only-svm > gpurun_out/r02b_ncu_persist_batch.log 2>&1
echo "ncu persist batch rc=$?"
ls -la gpurun_out/r02b_* | head -30
python tools/extract_reports.py r02b "see tools/call9.sh for the command of each report" -- gpurun_out/r02b_gram.ncu-rep gpurun_out/r02b_symtri.ncu-rep gpurun_out/r02b_lambda_batch.ncu-rep gpurun_out/r02b_persist.ncu-rep gpurun_out/r02b_persist_batch.ncu-rep > gpurun_out/r02b_ncu_summary.txt 2>&1
python tools/summarize_launches.py gpurun_out/r02b_launches.csv > gpurun_out/r02b_launches_summary.txt 2>&1
rm -f gpurun_out/r02b_lambda_batch.ncu-rep gpurun_out/r02b_persist.ncu-rep gpurun_out/r02b_persist_batch.ncu-rep gpurun_out/r02b_gram.ncu-rep
head -c 600000 gpurun_out/r02b_launches.csv > gpurun_out/r02b_launches_head.csv; rm -f gpurun_out/r02b_launches.csv
ls -la gpurun_out/r02b_*
