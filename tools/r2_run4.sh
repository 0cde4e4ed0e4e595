set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --batch 32 --streams 1 --no-cpu-baseline --no-e2e"
ICA_NO_GRAPH=1 ICA_NO_FUSE=1 $CMD > gpurun_out/r2_plain4.log 2>&1 && \
ICA_NO_GRAPH=1 ICA_NO_FUSE=1 ncu --set full --clock-control none --import-source on -k regex:ica_iterate_kernel -s 46 -c 3 -o gpurun_out/r2_prof4 $CMD > gpurun_out/r2_ncu4.log 2>&1
echo "ncu rc=$?"
