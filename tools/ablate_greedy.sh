#!/bin/bash
# Builds one library per timing-ablation mask of the persistent greedy kernel (I2L_ABL, decode_persistent.cu) into
# build_abl/ ; run on the GPU box with:  for m in 0 1 2 ...; do I2L_LIB=build_abl/libabl_$m.so python tools/time_greedy.py; done
set -e
cd "$(dirname "$0")/../hmer-img2latex_b200/csrc"
mkdir -p ../../build_abl
OBJS=$(ls build/*.o | grep -v decode_persistent.o)
for m in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -DI2L_ABL=$m -c decode_persistent.cu -o /tmp/dp_abl_$m.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build_abl/libabl_$m.so $OBJS /tmp/dp_abl_$m.o
done
