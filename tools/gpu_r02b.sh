#!/bin/bash
# round 2, session 2: A/B of the drain order (LC_FX_DRAIN) + MMA operands computed before the wait, parity, one ncu capture
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
