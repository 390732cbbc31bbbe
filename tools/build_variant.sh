#!/bin/bash
# tools/build_variant.sh <name> <log2n> <extra nvcc -D flags...>  -> scanner_b200/variants/lib_<name>.so (experiment builds)
set -e
NAME=$1; L2N=$2; shift 2
OUT=scanner_b200/variants; B=build/var_$NAME
mkdir -p $OUT $B
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -DSCN_ONLY_LOG2N=$L2N $@"
for f in scn_api scn_records scn_large scn_hackrf scn_convert scn_exchange scn_nccl scn_cluster scn_k_byte scn_k_short scn_k_shortc scn_k_float; do
  nvcc $FLAGS -c scanner_b200/csrc/$f.cu -o $B/$f.o &
done
wait
g++ -O2 -std=c++17 -fPIC -Iinclude -Iscanner_b200/csrc/host -c scanner_b200/csrc/host/frequencyTable.cpp -o $B/ft.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/lib_$NAME.so $B/*.o -cudart shared -lpthread -ldl
echo built $OUT/lib_$NAME.so
