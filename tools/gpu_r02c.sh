#!/bin/bash
# round 2, session 2: expander warps sleep between polls of b_empty (-DLC_FX_EXPANDER_SLEEP_NS=200 / 1000) instead of spinning
set -u
mkdir -p gpurun_out
bash tools/gpu_abn.sh ab_expander_sleep 50000 3 build/liblc_b200_sl200.so build/liblc_b200_sl1000.so
LC_B200_LIB=$PWD/build/liblc_b200_sl1000.so timeout 600 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_golden.py -x -q 2>&1 | tail -2 | tee -a gpurun_out/ab_expander_sleep.txt
LC_B200_LIB=$PWD/build/liblc_b200_sl1000.so python tools/loop_latency.py 2>&1 | tail -4 | tee -a gpurun_out/ab_expander_sleep.txt
python tools/loop_latency.py 2>&1 | tail -4 | tee -a gpurun_out/ab_expander_sleep.txt
