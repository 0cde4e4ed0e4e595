set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=3 > gpurun_out/r2_test11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_test11.log
tail -6 gpurun_out/r2_test11.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
$B > gpurun_out/r2_bench11_c2.json 2> gpurun_out/r2_bench11_c2.err; echo "c2 rc=$?"
ICA_NO_GATHER=1 $B > gpurun_out/r2_bench11_c2_nogather.json 2> gpurun_out/r2_bench11_c2_nogather.err
$B --batch 512 > gpurun_out/r2_bench11_c2_b512.json 2> gpurun_out/r2_bench11_c2_b512.err
$B --batch 512 --streams 8 > gpurun_out/r2_bench11_c2_b512_s8.json 2> gpurun_out/r2_bench11_c2_b512_s8.err
$B --streams 2 > gpurun_out/r2_bench11_c2_s2.json 2> gpurun_out/r2_bench11_c2_s2.err
$B --streams 8 > gpurun_out/r2_bench11_c2_s8.json 2> gpurun_out/r2_bench11_c2_s8.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench11_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], round(d['value']), round(d['ms_per_step'],2), round(r['frac'],4), round(r['kernel_ms_per_step'],2), round(r['pyramid_ms_per_step'],2))
    except Exception as e: print(f,'ERR',e)
PY
