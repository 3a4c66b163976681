#!/bin/bash
# A/B of two builds of the library on the C4-sized check (sustained load: the power cap decides): alternate them.
# The alternative build of profiles/r02_ab_zero_one_train.txt (both operands +1/-1, the round-1 encoding) is made here, on the CPU box, with
#   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -shared \
#        -DLC_FX_ZERO_ONE_TRAIN=0 -o build/liblc_b200_pm1.so slam_loop_closing_b200/csrc/lc_abi.cu
# and selected at run time with LC_B200_LIB (slam_loop_closing_b200/_native.py).
set -u
mkdir -p gpurun_out
OUT=gpurun_out/${1:-ab}.txt
ALT=${2:-build/liblc_b200_pm1.so}
: > $OUT
python -m pytest tests/test_gpu_tensor.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3 | tee -a $OUT
for i in 1 2 3; do
  echo "default build:" >> $OUT; python tools/tc_probe.py 0 100000 >> $OUT 2>&1
  echo "alternative build ($ALT):" >> $OUT; LC_B200_LIB=$PWD/$ALT python tools/tc_probe.py 0 100000 >> $OUT 2>&1
done
cat $OUT
