set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 2 4 8; do
  $TR --nproc-per-node $n --master-port 2950$n tools/h2d_ceiling.py > gpurun_out/r2_h2d_${n}gpu.json 2> gpurun_out/r2_h2d_${n}gpu.err
done
$TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench9_8gpu.json 2> gpurun_out/r2_bench9_8gpu.err; echo "bench8 rc=$?"
$TR --nproc-per-node 4 --master-port 29612 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2_bench9_4gpu.json 2> gpurun_out/r2_bench9_4gpu.err; echo "bench4 rc=$?"
for n in 2 4 8; do
  $TR --nproc-per-node $n --master-port 2962$n tools/bench_row_sharded.py --exchange peer > gpurun_out/r2_c5_${n}gpu_peer.json 2> gpurun_out/r2_c5_${n}gpu_peer.err; echo "c5 peer $n rc=$?"
done
$TR --nproc-per-node 8 --master-port 29631 tools/bench_row_sharded.py --exchange nccl > gpurun_out/r2_c5_8gpu_nccl.json 2> gpurun_out/r2_c5_8gpu_nccl.err; echo "c5 nccl 8 rc=$?"
python tools/bench_row_sharded.py --exchange peer > gpurun_out/r2_c5_1gpu_peer.json 2> gpurun_out/r2_c5_1gpu_peer.err
grep -h "^{" gpurun_out/r2_h2d_*gpu.json gpurun_out/r2_c5_*gpu_*.json | cut -c1-600
