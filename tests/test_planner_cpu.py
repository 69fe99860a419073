"""The time-segmentation planner (csrc/sos_plan.cpp choose_segmentation) through tfx_plan_segmentation: pure host arithmetic,
so its invariants are checked here on the CPU box over seeded random inputs, plus the cases the round-2 rules were tuned on."""
from __future__ import annotations

import ctypes

import numpy as np

from torchfx_b200 import _native

CAP = 61568  # streams one B200 holds in a wave of the channel-tile kernel (148 SMs x 13 warps x 32 lanes)


def plan(lanes, T, warm, cap=CAP, oversub=16):
    lib = _native.load()
    S, L, w = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
    lib.tfx_plan_segmentation(lanes, T, warm, cap, oversub, ctypes.byref(S), ctypes.byref(L), ctypes.byref(w))
    return S.value, L.value, w.value


def test_invariants_on_random_inputs():
    rng = np.random.default_rng(2)
    for it in range(20000):
        C = int(rng.integers(1, 4097))
        if it % 2:
            C = (C + 31) // 32 * 32
        T = int(rng.integers(1, 40_000_000 if it % 3 == 0 else 300_000))
        warm = int(rng.integers(0, 50_000 if it % 5 == 0 else 3_000))
        cap = CAP if it % 2 else int(rng.integers(32, 200_000))
        ov = 16 if it % 4 < 2 else 1
        S, L, w = plan(C, T, warm, cap, ov)
        assert S >= 1
        if S == 1:
            assert L == T and w == 0
            continue
        assert L % 64 == 0 and w % 64 == 0 and w >= warm             # aligned starts; the warm-up covers what was asked
        assert (S - 1) * L < T <= S * L and T - (S - 1) * L >= 2     # the segments tile [0, T), the last one keeps a 2-sample tail
        assert L >= w and L >= 512                                   # a segment is at least its warm-up
        assert C * S <= cap * ov + C                                 # never more items than the (oversubscribed) wave


def test_no_split_when_the_filter_never_forgets_or_the_signal_is_short():
    assert plan(1024, 28_800_000, -1)[0] == 1
    assert plan(1024, 600, 256)[0] == 1
    assert plan(100_000, 28_800_000, 256)[0] == 1  # more channels than the wave holds: nothing to gain from time splits


def test_round2_rules():
    # config 2 (1024 ch x 10 min, warm-up 384): 16 items per resident warp
    S, L, _ = plan(1024, 28_800_000, 384)
    assert S == 960 and L == 30016
    # config 4's mixed-precision chain (warm-up 1728 samples) on the 1 / 2 / 4 / 8-GPU shards: the 1/32 rule alone would leave
    # one item per warp on the shards; 1.73 items per warp as long as the warm-up stays <= 1/4 of a segment
    for C, want in ((2048, 52), (1024, 105), (512, 209), (256, 413)):
        S, L, w = plan(C, 2_880_000, 1728)
        assert S == want and w == 1728 and w * 4 <= L + 64, (C, S, L)
        assert 1.7 <= C * S / CAP <= 1.8
    # the 8-branch SUM bank (warm-up 4480): its 512-channel shard cannot reach 1.73 items within that limit -> one item per warp,
    # not a partial stagger (measured slower)
    assert plan(512, 2_880_000, 4480)[0] == CAP // 512
    S, _, _ = plan(1024, 2_880_000, 4480)
    assert 1.7 <= 1024 * S / CAP <= 1.8


def test_fir_auto_rule_mirrored_in_python():
    """filter/fir.py decides whether to build an overlap-save plan with the same rule csrc/fir.cu's TFX_FIR_AUTO uses;
    tfx_fir_workspace_bytes is host arithmetic (0 bytes <=> the direct form runs), so the two are compared here on a grid."""
    from torchfx_b200.filter.fir import _auto_takes_overlap_save

    lib = _native.load()
    for K in (1, 8, 32, 56, 57, 64, 96, 97, 300, 1024, 1025, 70000):
        for C, T in ((1, 1000), (2, 48000), (8, 480000), (64, 262144), (64, 262145), (256, 2880000)):
            ols = lib.tfx_fir_workspace_bytes(C, T, K, _native.TFX_FIR_AUTO) > 0
            assert ols == _auto_takes_overlap_save(K, C * T), (K, C, T)
    # forced modes: the direct form is kept up to 1024 taps, beyond that it is overlap-save whatever was asked
    assert lib.tfx_fir_workspace_bytes(4, 10000, 1024, _native.TFX_FIR_DIRECT) == 0
    assert lib.tfx_fir_workspace_bytes(4, 10000, 1025, _native.TFX_FIR_DIRECT) > 0
    assert lib.tfx_fir_workspace_bytes(4, 10000, 8, _native.TFX_FIR_OLS) > 0
    # the plan holds the twiddle tables and one 128 KB spectrum row per 8192-tap partition
    assert lib.tfx_fir_plan_bytes(65536) - lib.tfx_fir_plan_bytes(8192) == 7 * 8192 * 16
