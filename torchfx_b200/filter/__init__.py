"""Filter classes -- same 33 exported names as the reference (src/torchfx/filter/__init__.py:38-72)."""
from .biquad import Biquad, BiquadAllPass, BiquadBPF, BiquadBPFPeak, BiquadHPF, BiquadLPF, BiquadNotch
from .filterbank import LogFilterBank
from .fir import FIR, DesignableFIR
from .fused import FusedSOSCascade
from .iir import (
    IIR,
    AllPass,
    Butterworth,
    Chebyshev1,
    Chebyshev2,
    Elliptic,
    HiButterworth,
    HiChebyshev1,
    HiChebyshev2,
    HiElliptic,
    HiLinkwitzRiley,
    HiShelving,
    LinkwitzRiley,
    LoButterworth,
    LoChebyshev1,
    LoChebyshev2,
    LoElliptic,
    LoLinkwitzRiley,
    LoShelving,
    Notch,
    ParametricEQ,
    Peaking,
)
from ._base import AbstractFilter, ParallelFilterCombination

__all__ = [
    "AllPass", "Biquad", "BiquadAllPass", "BiquadBPF", "BiquadBPFPeak", "BiquadHPF", "BiquadLPF", "BiquadNotch",
    "Butterworth", "Chebyshev1", "Chebyshev2", "DesignableFIR", "Elliptic", "FIR", "FusedSOSCascade",
    "HiButterworth", "HiChebyshev1", "HiChebyshev2", "HiElliptic", "HiLinkwitzRiley", "HiShelving", "IIR",
    "LinkwitzRiley", "LoButterworth", "LoChebyshev1", "LoChebyshev2", "LoElliptic", "LoLinkwitzRiley",
    "LogFilterBank", "LoShelving", "Notch", "ParametricEQ", "Peaking",
]
