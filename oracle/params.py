"""Ambiguity knobs of SURVEY.md Appendix B, shared by both oracle statements (test infrastructure)."""
from dataclasses import dataclass


@dataclass
class Params:
    loop_threshold: float = 0.7   # include/loop_closing.hpp:31 (README.md:108 passes 0.15)
    min_loop_gap: int = 30        # include/loop_closing.hpp:31, README.md:109
    min_good: int = 50            # README.md:124
    filter_floor: int = 0         # B2
    filter_strict: bool = False   # B1
    min_init: int = -1            # B3 (<0: true minimum)
    gap_strict: bool = False      # B4
    thr_inclusive: bool = False   # B5
