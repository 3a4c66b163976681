#!/bin/bash
# round 2, session 2: epilogue specialised for blocks of all eight query tiles (-DLC_FX_ALL8=1)
set -u
mkdir -p gpurun_out
bash tools/gpu_abn.sh ab_all8 50000 3 build/liblc_b200_all8.so
LC_B200_LIB=$PWD/build/liblc_b200_all8.so timeout 600 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_golden.py -x -q 2>&1 | tail -2 | tee -a gpurun_out/ab_all8.txt
