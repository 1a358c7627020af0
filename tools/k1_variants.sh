#!/bin/bash
# Build K1 register-tile variants of libsober_b200.so into gpurun_out/variants/ (same ABI), for tools/k1_time.py
set -e
cd "$(dirname "$0")/../sober_b200/csrc"
OUT=../../k1_variants; mkdir -p $OUT
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --fmad=true"
for v in "4 2 1" "4 2 3" "4 4 1" "2 4 2" "2 4 3" "2 2 3" "4 1 2" "8 1 1"; do
  set -- $v
  $NV -DSOBER_REC_TL=$1 -DSOBER_REC_TG=$2 -DSOBER_REC_MINB=$3 -Xptxas -v -c group_accumulate.cu -o /tmp/ga_$1_$2_$3.o 2> /tmp/ga_$1_$2_$3.log
  grep -A2 "group_records_kernelILi6ELi3E" /tmp/ga_$1_$2_$3.log | grep -E "Used" | head -1 | sed "s/^/TL=$1 TG=$2 MINB=$3: /"
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/lib_$1_$2_$3.so util.o /tmp/ga_$1_$2_$3.o car_eliminate.o car_cluster.o stream_ops.o -lcudart
done
