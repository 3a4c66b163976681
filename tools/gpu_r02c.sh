#!/bin/bash
# round 2, session 2: one TMEM load in flight per warp, each load's address depending on the previous load's data (-DLC_FX_DRAIN_CHAIN=1)
set -u
mkdir -p gpurun_out
bash tools/gpu_abn.sh ab_chain 50000 3 build/liblc_b200_chain.so
LC_B200_LIB=$PWD/build/liblc_b200_chain.so timeout 600 python -m pytest tests/test_gpu_tensor.py -x -q 2>&1 | tail -2 | tee -a gpurun_out/ab_chain.txt
