#!/bin/bash
# Builds a tuning variant of the library next to the real one: tools/build_variant.sh NAME "-DICA_BH=36 ..."
# -> inverse_compositional_algorithm_b200/variants/libica_NAME.so ; run with ICA_LIB_PATH=<that file>.
set -e
cd "$(dirname "$0")/.."
name=$1; flags=$2
out=inverse_compositional_algorithm_b200/variants
mkdir -p $out /tmp/ica_var_$name
objs=""
for f in ica_iterate ica_march ica_pyramid ica_capi ica_helpers ica_generate; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden $flags \
    -c inverse_compositional_algorithm_b200/csrc/$f.cu -o /tmp/ica_var_$name/$f.o &
  objs="$objs /tmp/ica_var_$name/$f.o"
done
wait
nvcc -shared -o $out/libica_$name.so $objs -gencode arch=compute_100a,code=sm_100a
echo $out/libica_$name.so
