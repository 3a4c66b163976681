#!/bin/bash
# round 2, session 2: what does the drain cost the tensor pipe?  Timing-only builds (wrong results, -DLC_FX_DRAIN_EXPERIMENT=1/2/3):
# x1 = all TMEM loads, no fold; x2 = 64 of 112 columns loaded and folded; x3 = all loads, 64 columns folded
set -u
mkdir -p gpurun_out
bash tools/gpu_abn.sh ab_drain_cost 50000 2 build/liblc_b200_old.so build/liblc_b200_x1.so build/liblc_b200_x2.so build/liblc_b200_x3.so
