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
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
for v in "13 26" "15 26"; do
  set -- $v
  ICA_NVCC_EXTRA="-DICA_CONSUMER_WARPS=$1 -DICA_BH=$2" python -m inverse_compositional_algorithm_b200.build --force > /dev/null 2>&1
  ICA_NVCC_EXTRA="-DICA_CONSUMER_WARPS=$1 -DICA_BH=$2" $B > gpurun_out/r2_b8_w$1_bh$2.json 2> gpurun_out/r2_b8_w$1_bh$2.err
  echo "variant $v rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_b8_*.json'))+sorted(glob.glob('gpurun_out/r2_bench8_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d.get('roofline',{})
        print(f, round(d['value'],2), round(d['ms_per_step'],2), r.get('frac'), r.get('kernel_ms_per_step'), (d.get('e2e') or {}).get('value'), (d.get('e2e_full_tuple') or {}).get('value'))
    except Exception as e: print(f,'ERR',e)
PY
