"""GPU tier: a short run of the randomised parity soak (tests/fuzz_gpu.py) - random shapes, kernels and knobs through the
C ABI against the CPU oracle, bit-exact.  `python tests/fuzz_gpu.py 240` runs it for as long as one likes."""
import pytest

pytestmark = pytest.mark.gpu


def test_random_cases_are_bit_exact():
    import fuzz_gpu
    n, _ = fuzz_gpu.soak(15.0, 12345)
    assert n >= 20
