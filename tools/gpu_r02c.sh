#!/bin/bash
# round 2, session 2: the third / fourth TMEM load of a drain issued earlier (-DLC_FX_DRAIN_ORDER=1 [-DLC_FX_DRAIN_ORDER2=2|3])
set -u
mkdir -p gpurun_out
bash tools/gpu_abn.sh ab_drain_order 50000 3 build/liblc_b200_old.so build/liblc_b200_o1.so build/liblc_b200_o4.so build/liblc_b200_o5.so
for L in build/liblc_b200_o5.so; do
  LC_B200_LIB=$PWD/$L timeout 600 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_golden.py -x -q 2>&1 | tail -2 | tee -a gpurun_out/ab_drain_order.txt
done
