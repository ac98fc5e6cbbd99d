#!/bin/bash
# Builds mcevidence_b200/variants/lib_<name>.so with extra -D flags for knn_filter_tc.cu (development A/B runs):
#   tools/build_variant.sh coop_hoist "-DTC_HIT_COOP=1 -DTC_ISSUE_HOIST=1"
set -e
name=$1; flags=$2
cd "$(dirname "$0")/../mcevidence_b200/csrc"
mkdir -p ../variants build_var
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-O2,-Wall $flags -c knn_filter_tc.cu -o build_var/tc_$name.o
objs=$(ls build/*.o | grep -v knn_filter_tc.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/lib_$name.so $objs build_var/tc_$name.o -Xlinker -lpthread -Xlinker -ldl
echo built ../variants/lib_$name.so
