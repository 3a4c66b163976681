#!/bin/bash
# alternate the default build with several alternatives: gpu_abn.sh TAG K reps lib1 lib2 ...
set -u
OUT=gpurun_out/$1.txt; K=$2; REPS=$3; shift 3
: > $OUT
for i in $(seq $REPS); do
  echo -n "default : " >> $OUT; python tools/tc_probe.py 0 $K >> $OUT 2>&1
  for L in "$@"; do echo -n "$L : " >> $OUT; LC_B200_LIB=$PWD/$L python tools/tc_probe.py 0 $K >> $OUT 2>&1; done
done
cat $OUT
