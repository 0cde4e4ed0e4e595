set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2_test10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_test10.log
tail -6 gpurun_out/r2_test10.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench10_c2.json 2> gpurun_out/r2_bench10_c2.err; echo "c2 rc=$?"
CMD="python bench.py --workload c3 --steps 1 --warmup 1 --batch 64 --streams 1 --no-cpu-baseline --no-e2e"
ICA_NO_GRAPH=1 $CMD > gpurun_out/r2_plain10.log 2>&1 && \
ICA_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:ica_iterate_kernel -s 40 -c 3 -o gpurun_out/r2_prof10_gray $CMD > gpurun_out/r2_ncu10.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench10_c2.json').read().strip().splitlines()[-1]); r=d['roofline']
print('c2', round(d['value']), round(d['ms_per_step'],2), round(r['frac'],4), round(r['kernel_ms_per_step'],2), round(r['pyramid_ms_per_step'],2))
PY
