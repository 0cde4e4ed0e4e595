set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r2_bench8_c2.json 2> gpurun_out/r2_bench8_c2.err; echo "c2 rc=$?"
python bench.py --workload c3 --no-cpu-baseline > gpurun_out/r2_bench8_c3.json 2> gpurun_out/r2_bench8_c3.err; echo "c3 rc=$?"
python bench.py --workload c3 --batch 2048 --no-cpu-baseline --no-full-tuple > gpurun_out/r2_bench8_c3_b2048.json 2> gpurun_out/r2_bench8_c3_b2048.err; echo "c3 2048 rc=$?"
python bench.py --workload c4 --no-cpu-baseline > gpurun_out/r2_bench8_c4.json 2> gpurun_out/r2_bench8_c4.err; echo "c4 rc=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2_bench8_ref.json 2> gpurun_out/r2_bench8_ref.err; echo "ref rc=$?"
python tools/h2d_ceiling.py > gpurun_out/r2_h2d_1gpu.json 2>&1
python tools/latency_probe.py > gpurun_out/r2_latency8.log 2>&1
tail -3 gpurun_out/r2_bench8_c2.err
