"""``FilterChain``: the ``nn.Sequential`` produced by ``f1 | f2`` (reference src/torchfx/chain.py:31-67)."""
from __future__ import annotations

from torch import nn


class FilterChain(nn.Sequential):
    """Sequential container that flattens nested chains, so ``(f1 | f2) | f3`` holds three
    steps and ``Wave._materialize`` sees one run of IIR filters to fuse (chain.py:52-59)."""

    def __init__(self, *modules: nn.Module) -> None:
        steps: list[nn.Module] = []
        for m in modules:
            steps.extend(m.children() if isinstance(m, FilterChain) else [m])
        super().__init__(*steps)

    def __or__(self, other: nn.Module) -> "FilterChain":
        if not isinstance(other, nn.Module):
            return NotImplemented
        return FilterChain(*self.children(), other)

    def __ror__(self, other: object) -> "FilterChain":
        return NotImplemented
