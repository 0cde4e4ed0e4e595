set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2_test6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_test6.log
tail -8 gpurun_out/r2_test6.log
for ex in nccl peer; do
python tools/bench_row_sharded.py --exchange $ex > gpurun_out/r2_c5_1gpu_$ex.json 2> gpurun_out/r2_c5_1gpu_$ex.err; echo "c5 1gpu $ex rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/bench_row_sharded.py --exchange $ex > gpurun_out/r2_c5_2gpu_$ex.json 2> gpurun_out/r2_c5_2gpu_$ex.err; echo "c5 2gpu $ex rc=$?"
done
cat gpurun_out/r2_c5_*.json
tail -3 gpurun_out/r2_c5_2gpu_peer.err
