#!/bin/bash
# alternate two builds on short (K=4000) and one-launch-long (K=50000) checks
set -u
OUT=gpurun_out/${1:-ab2}.txt
ALT=${2:-build/liblc_b200_pm1.so}
: > $OUT
for K in 4000 50000; do
for i in 1 2 3; do
  echo -n "default     : " >> $OUT; python tools/tc_probe.py 0 $K >> $OUT 2>&1
  echo -n "alternative : " >> $OUT; LC_B200_LIB=$PWD/$ALT python tools/tc_probe.py 0 $K >> $OUT 2>&1
done
done
cat $OUT
