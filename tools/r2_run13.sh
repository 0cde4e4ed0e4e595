set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 150 --durations=3 > gpurun_out/r2_test13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_test13.log
tail -5 gpurun_out/r2_test13.log
timeout 600 python bench.py > gpurun_out/r2_bench13_c2.json 2> gpurun_out/r2_bench13_c2.err; echo "c2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench13_c2.json').read().strip().splitlines()[-1]); r=d['roofline']
print(round(d['value']), round(d['ms_per_step'],2), round(r['frac'],4), round(r['kernel_ms_per_step'],2), round(r['pyramid_ms_per_step'],2), d['e2e']['value'], d['e2e_full_tuple']['value'], d.get('cpu_baseline'))
PY
