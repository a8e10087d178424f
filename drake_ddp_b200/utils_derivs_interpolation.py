"""Keypoint configuration records, same names and field order as the reference's
utils_derivs_interpolation.py:3-14 so scripts can construct them positionally
(e.g. acrobot.py:115)."""
from dataclasses import dataclass


@dataclass
class derivs_interpolation:
    keypoint_method: str                 # 'setInterval' | 'adaptiveJerk' | 'iterativeError'
    minN: int
    maxN: int
    jerk_threshold: float
    iterative_error_threshold: float


@dataclass
class index_tuple:
    start_index: int
    end_index: int
