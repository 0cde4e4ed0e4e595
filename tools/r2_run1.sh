set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=8 -s > gpurun_out/r2_test1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_test1.log
tail -40 gpurun_out/r2_test1.log
