set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2_test3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_test3.log
tail -15 gpurun_out/r2_test3.log
python __graft_entry__.py --smoke > gpurun_out/r2_smoke3.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; echo "bench rc=$?"
ICA_NO_FUSE=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench3_nofuse.json 2> gpurun_out/r2_bench3_nofuse.err; echo "bench rc=$?"
python tools/latency_probe.py > gpurun_out/r2_latency3.log 2>&1
ICA_NO_FUSE=1 python tools/latency_probe.py > gpurun_out/r2_latency3_nofuse.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload c3 > gpurun_out/r2_bench3_c3.json 2> gpurun_out/r2_bench3_c3.err; echo "bench c3 rc=$?"
