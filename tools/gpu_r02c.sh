#!/bin/bash
# round 2, session 2: timing-only builds (wrong results, -DLC_FX_DRAIN_EXPERIMENT=n) that take the drain's cost apart:
# x1 = all TMEM loads at once, no fold; x2 = 64 of 112 columns loaded and folded; x3 = all loads, 64 columns folded;
# x4 = x1 through the 16x256b access shape; x6 = no loads, 56 FMNMX3 on opaque values
set -u
mkdir -p gpurun_out
bash tools/gpu_abn.sh ab_drain_cost2 50000 2 build/liblc_b200_x4.so build/liblc_b200_x6.so
