#!/bin/bash
# usage: tools/run_variants.sh "<bench args>" variant...   (variants built by tools/build_variant.sh)
args=$1; shift
for v in "$@"; do
  ICA_LIB_PATH=$PWD/inverse_compositional_algorithm_b200/variants/libica_$v.so timeout 300 python bench.py $args > gpurun_out/var_$v.log 2>&1
  echo "$v: $(grep -o '"value": [0-9.]*' gpurun_out/var_$v.log | head -1) $(grep -o '"frac": [0-9.]*' gpurun_out/var_$v.log) $(grep -o '"kernel_ms_per_step": [0-9.]*' gpurun_out/var_$v.log) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/var_$v.log) $(tail -1 gpurun_out/var_$v.log | cut -c1-60)"
done
