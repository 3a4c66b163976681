#!/bin/bash
# round 2, session 2: A/B of the drain order (LC_FX_DRAIN) + MMA operands computed before the wait, parity, one ncu capture
# (Record of a session-2 A/B run.  The alternative builds were made on the CPU box from working-tree variants of lc_mxf4.cuh,
#  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -shared -D<macro> \
#       -o build/liblc_b200_<tag>.so slam_loop_closing_b200/csrc/lc_abi.cu
#  and selected with LC_B200_LIB; what each variant was and what it measured is in profiles/r02_ab_drain_schedule.txt.  Of the
#  macros only LC_FX_DRAIN_EXPERIMENT (timing-only builds x1-x6) is still in the source: the rejected variants were not kept.)
set -u
mkdir -p gpurun_out
echo "== parity default" | tee gpurun_out/ab_drain_parity.txt
timeout 900 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_golden.py -x -q 2>&1 | tail -3 | tee -a gpurun_out/ab_drain_parity.txt
LIBS="build/liblc_b200_old.so build/liblc_b200_d0m1.so build/liblc_b200_d2m1.so"
bash tools/gpu_abn.sh ab_drain 50000 3 $LIBS
python tools/tc_probe.py 0 4000 > gpurun_out/r02g_probe.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mxf4_match -s 3 -c 1 -f -o gpurun_out/r02g_mxf4 \
  python tools/tc_probe.py 0 4000 > gpurun_out/r02g_ncu_full.log 2>&1
echo "full capture rc=$?"
