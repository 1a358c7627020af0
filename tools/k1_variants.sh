#!/bin/bash
# Builds K1 variants of libsober_b200.so into k1_variants/ (same ABI; shipped to the GPU box, git-ignored) for
# tools/k1_time.py:  SOBER_B200_LIB=k1_variants/lib_<name>.so python tools/k1_time.py
# name = TL_TG_MINB_ROWS_TABBITS_DEG_INTHALF
cd "$(dirname "$0")/../sober_b200/csrc"
make -j8 > /dev/null
OUT=../../k1_variants; mkdir -p $OUT; rm -f $OUT/*.so
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --fmad=true"
OTHERS=$(ls *.o | grep -v group_accumulate.o)
for v in "2 4 2 16 6 5 0" "2 4 2 16 8 3 1" "2 4 2 16 11 2 1" "2 4 2 32 8 3 1" "4 4 1 16 8 3 1" "4 2 2 16 11 2 1"; do
  set -- $v
  name=$1_$2_$3_$4_$5_$6_$7
  $NV -DSOBER_REC_TL=$1 -DSOBER_REC_TG=$2 -DSOBER_REC_MINB=$3 -DSOBER_REC_ROWS=$4 -DSOBER_EXP_TAB_BITS=$5 -DSOBER_EXP_DEG=$6 \
      -DSOBER_SQRT_INTHALF=$7 -Xptxas -v -c group_accumulate.cu -o /tmp/ga_$name.o 2> /tmp/ga_$name.log || { echo "$name: does not compile"; continue; }
  grep -A2 "group_records_kernelILi6ELi3ELi$1ELi$2ELb0E" /tmp/ga_$name.log | grep -E "Used" | head -1 | sed "s/^/$name: /"
  grep -A1 "group_records_kernelILi6ELi3ELi$1ELi$2ELb0E" /tmp/ga_$name.log | grep -E "spill" | head -1
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/lib_$name.so $OTHERS /tmp/ga_$name.o -lcudart
done
