#!/bin/bash
# A/B of two builds of the library on the C4-sized check (sustained load: the power cap decides): alternate them
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
