set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=5 -s > gpurun_out/r2_test7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_test7.log
grep -E "log line|passed|failed|FAILED|Error" gpurun_out/r2_test7.log | tail -20
for ex in nccl peer; do
python tools/bench_row_sharded.py --exchange $ex > gpurun_out/r2_c5_1gpu_$ex.json 2> gpurun_out/r2_c5_1gpu_$ex.err; echo "c5 1gpu $ex rc=$?"
done
cat gpurun_out/r2_c5_1gpu_*.json
