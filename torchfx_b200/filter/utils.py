"""Reference: src/torchfx/filter/utils.py."""
from ..typing import FilterOrderScale


def compute_order(o: int, scale: FilterOrderScale) -> int:
    if scale == "db":
        return o // 6
    if scale == "linear":
        return o
    raise ValueError(f"unknown order scale {scale!r}")
