set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2_test12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_test12.log
tail -3 gpurun_out/r2_test12.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
$B > gpurun_out/r2_bench12_c2.json 2> gpurun_out/r2_bench12_c2.err; echo "c2 rc=$?"
$B --batch 512 > gpurun_out/r2_bench12_c2_b512.json 2> gpurun_out/r2_bench12_c2_b512.err
$B --batch 1024 > gpurun_out/r2_bench12_c2_b1024.json 2> gpurun_out/r2_bench12_c2_b1024.err
$B --batch 1024 --streams 8 > gpurun_out/r2_bench12_c2_b1024_s8.json 2> gpurun_out/r2_bench12_c2_b1024_s8.err
$B --batch 2048 --steps 3 > gpurun_out/r2_bench12_c2_b2048.json 2> gpurun_out/r2_bench12_c2_b2048.err
python tools/latency_probe.py > gpurun_out/r2_latency12.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench12_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], round(d['value']), round(d['ms_per_step'],2), round(r['frac'],4), round(r['kernel_ms_per_step'],2), round(r['pyramid_ms_per_step'],2))
    except Exception as e: print(f,'ERR',e)
PY
cat gpurun_out/r2_latency12.log
