"""CPU oracle for the Hamming hot path - TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package; the product (slam_loop_closing_b200) never does and has no CPU fallback.

Two independent statements of the same algorithm live here:
  * c_oracle  - ctypes binding of lc_oracle.c, the plain-C restatement;
  * cv2_ref   - the real OpenCV matcher (cv2.BFMatcher(NORM_HAMMING), the call behind the reference's
                `matcher_`, include/loop_closing.hpp:73) plus README.md:117-126's rules in numpy.
The C restatement is pinned against cv2_ref in tests/test_oracle.py and tests/golden/.
"""
