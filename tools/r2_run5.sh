set -x
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
for v in "30 2" "22 2" "19 3" "24 2"; do
  set -- $v
  ICA_NVCC_EXTRA="-DICA_BH=$1 -DICA_STAGES_RGB=$2" python -m inverse_compositional_algorithm_b200.build --force > /dev/null 2>&1
  ICA_NVCC_EXTRA="-DICA_BH=$1 -DICA_STAGES_RGB=$2" $B > gpurun_out/r2_b5_bh$1_s$2.json 2> gpurun_out/r2_b5_bh$1_s$2.err
  echo "variant $v rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_b5_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f, round(d['value']), round(d['ms_per_step'],2), round(r['frac'],4), round(r['kernel_ms_per_step'],2))
    except Exception as e: print(f,'ERR',e)
PY
# the generator kernel against its numpy mirror + a registration of generated pairs
python - <<'PY' > gpurun_out/r2_gen5.log 2>&1
import numpy as np, torch, time
from inverse_compositional_algorithm_b200 import synthetic
from inverse_compositional_algorithm_b200.transformation import TransformType, end_point_error
from inverse_compositional_algorithm_b200.inverse_compositional_algorithm import register_batch_device
t=TransformType.HOMOGRAPHY
for (H,W,C,occ) in ((96,128,3,0.2),(120,160,1,0.0)):
    I1,I2,p=synthetic.make_batch_device(4,H,W,C,t,seed=3,pair_offset=2,occlusion=occ,margin=32)
    a1,a2,pg=synthetic.make_pair_hash(3,3,H,W,C,t,occlusion=occ,margin=32)
    d1=np.abs(I1[1].cpu().numpy()-a1); d2=np.abs(I2[1].cpu().numpy()-a2)
    print(H,W,C,'I1 diff max',d1.max(),'frac',(d1>0).mean(),'I2 diff max',d2.max(),'frac',(d2>0).mean(), 'p equal', np.allclose(p[1,:8],pg))
torch.cuda.synchronize(); t0=time.time()
I1,I2,p=synthetic.make_batch_device(64,1024,1024,3,t,seed=1)
torch.cuda.synchronize(); print('64 pairs 1024^2 RGB generated in',time.time()-t0,'s')
pr,err,it=register_batch_device(I1,I2,t,nscales=5,robust_type=3,delta=10)
pr=pr.cpu().numpy()
e=[end_point_error(pr[i],p[i],t,1024,1024)[1] for i in range(64)]
print('EPE vs ground truth: median',np.median(e),'max',np.max(e),'iters mean',it.sum(1).mean())
PY
cat gpurun_out/r2_gen5.log
