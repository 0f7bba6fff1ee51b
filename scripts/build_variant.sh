#!/bin/bash
# Kernel A/B runs: build another copy of the library with extra nvcc flags for the Nn=10 instantiation.
#   scripts/build_variant.sh <name> "<flags>"   ->  fitoct_b200/variants/lib_<name>.so   (load with FOCT_LIB_PATH)
# Only the Nn = 0 and Nn = 10 translation units are linked (the others resolve to stubs), no -lineinfo: ~10 MB per variant.
set -e
cd "$(dirname "$0")/../fitoct_b200/csrc"
name=$1; flags=$2
out=../variants; mkdir -p $out build/v_$name
ARCH="-gencode arch=compute_100a,code=sm_100a"
for nn in 0 10; do
  nvcc $ARCH -O3 -std=c++17 -Xcompiler -fPIC -DFOCT_INST_NN=$nn $flags -c foct_inst.cu -o build/v_$name/inst_$nn.o &
done
nvcc $ARCH -O3 -std=c++17 -Xcompiler -fPIC -DFOCT_VARIANT_NNS $flags -c foct_lib.cu -o build/v_$name/foct_lib.o &
nvcc $ARCH -O3 -std=c++17 -Xcompiler -fPIC $flags -c foct_prep.cu -o build/v_$name/foct_prep.o &
wait
nvcc $ARCH -shared -o $out/lib_$name.so build/v_$name/*.o -lcudart
ls -la $out/lib_$name.so
