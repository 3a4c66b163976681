"""One short run of the library's pure tcgen05.mma kind::mxf4 issue loop (lc_probe_mxf4_peak), for ncu:
  ncu -k regex:mxf4_peak -c 1 --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,... python tools/peak_probe_once.py
(round 2: tensor pipe 99.5 % active, sm__mem_tensor_cycles_active 99.5 %, shared-memory operand pipe 73 % at N = 256)."""
import sys
sys.path.insert(0, ".")
import slam_loop_closing_b200 as lc

with lc.Context() as c:
    print(c.probe_mxf4_peak(5))
