set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --batch 32 --streams 1 --no-cpu-baseline --no-e2e"
ICA_NO_GRAPH=1 $CMD > gpurun_out/r2_plain_final.json 2> gpurun_out/r2_plain_final.err && \
ICA_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:ica_|pyr_|minmax|clip_' -c 700 --csv --log-file gpurun_out/r2_launches_final.csv $CMD > gpurun_out/r2_ncu14a.log 2>&1
echo "ncu list rc=$?"
ICA_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:ica_iterate_kernel -s 46 -c 3 -o gpurun_out/r2_prof_final $CMD > gpurun_out/r2_ncu14b.log 2>&1
echo "ncu full rc=$?"
ICA_NO_GRAPH=1 ncu --set full --clock-control none -k 'regex:pyr_fused_fast|ica_solve_kernel' -s 4 -c 6 -o gpurun_out/r2_prof_final_other $CMD > gpurun_out/r2_ncu14c.log 2>&1
echo "ncu other rc=$?"
