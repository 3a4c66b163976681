import sys
sys.path.insert(0, ".")
import slam_loop_closing_b200 as lc
with lc.Context() as c:
    print(c.probe_mxf4_peak(5))
