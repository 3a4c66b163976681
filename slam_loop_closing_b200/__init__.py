"""slam_loop_closing_b200 - B200-native brute-force Hamming matching for SLAM loop closing.

The product is csrc/liblc_b200.so (CUDA sm_100a kernels behind the C ABI of include/lc_abi.h).  This
package is the thin host side: the ctypes binding, the mirror of the reference's LoopClosingSystem
interface and the torch.distributed plumbing for keyframe-sharded multi-GPU loop checks.
"""
from ._native import (Config, Context, KF_RESULT, LOOP_CANDIDATE, LcError, KERNEL_AUTO, KERNEL_CSA,  # noqa: F401
                      KERNEL_MXF4, KERNEL_POPC, KERNEL_TENSOR, MATCH, UINT64_MAX, default_config, library_path, load)
from .loop_closing import B200Matcher, DMatch, Frame, LoopCandidate, LoopClosingSystem  # noqa: F401

__all__ = ["Config", "Context", "KF_RESULT", "LOOP_CANDIDATE", "LcError", "default_config", "library_path", "load",
           "B200Matcher", "DMatch", "Frame", "LoopCandidate", "LoopClosingSystem",
           "KERNEL_AUTO", "KERNEL_CSA", "KERNEL_POPC", "KERNEL_TENSOR", "KERNEL_MXF4", "MATCH", "UINT64_MAX"]
