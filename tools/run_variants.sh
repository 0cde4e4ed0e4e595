for v in base b1w23s3 b1w23s3h34 b1w19s3 b1w15s4; do
  ICA_LIB_PATH=$PWD/inverse_compositional_algorithm_b200/variants/libica_$v.so timeout 200 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/var_$v.log 2>&1
  echo "$v: $(grep -o '"value": [0-9.]*' gpurun_out/var_$v.log | head -1) $(grep -o '"frac": [0-9.]*' gpurun_out/var_$v.log) $(grep -o '"kernel_ms_per_step": [0-9.]*' gpurun_out/var_$v.log) $(tail -1 gpurun_out/var_$v.log | cut -c1-80)"
done
