#!/bin/bash
# round 2, session 2: seven 16-column TMEM loads per drain instead of 32 + 32 + 32 + 16 (-DLC_FX_DRAIN_X16=1)
set -u
mkdir -p gpurun_out
bash tools/gpu_abn.sh ab_x16 50000 3 build/liblc_b200_x16.so
LC_B200_LIB=$PWD/build/liblc_b200_x16.so timeout 600 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_golden.py -x -q 2>&1 | tail -2 | tee -a gpurun_out/ab_x16.txt
