#!/bin/bash
# round 2, session 2: one TMEM load in flight per warp, each load's address depending on the previous load's data (-DLC_FX_DRAIN_CHAIN=1)
# (Record of a session-2 A/B run.  The alternative builds were made on the CPU box from working-tree variants of lc_mxf4.cuh,
#  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -shared -D<macro> \
#       -o build/liblc_b200_<tag>.so slam_loop_closing_b200/csrc/lc_abi.cu
#  and selected with LC_B200_LIB; what each variant was and what it measured is in profiles/r02_ab_drain_schedule.txt.  Of the
#  macros only LC_FX_DRAIN_EXPERIMENT (timing-only builds x1-x6) is still in the source: the rejected variants were not kept.)
set -u
mkdir -p gpurun_out
bash tools/gpu_abn.sh ab_chain 50000 3 build/liblc_b200_chain.so
LC_B200_LIB=$PWD/build/liblc_b200_chain.so timeout 600 python -m pytest tests/test_gpu_tensor.py -x -q 2>&1 | tail -2 | tee -a gpurun_out/ab_chain.txt
