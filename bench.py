#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native filter engine.

Metric (BASELINE.json): Gsamples/s of a 1024-channel x 10-minute @48 kHz float32 stream
through a 4-section SOS cascade (LoButterworth(5 kHz, order 8)), at 1/2/4/8 B200, next to
the HBM roofline and the reference's CPU path.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's own CPU kernel, same metric

One "step" = one pass of the whole workload through the fused cascade (one C-ABI call:
warm-up launch + main launch).  The 1024 channels are sharded over the ranks (strong
scaling: the job is fixed, BASELINE names 1024 channels), no collective on the data path.
Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = 48000
CHANNELS = 1024
SECONDS = float(os.environ.get("TFX_BENCH_SECONDS", "600"))  # override only for local experiments
SECTIONS = 4
E2E_SECONDS = float(os.environ.get("TFX_BENCH_E2E_SECONDS", "60"))
CPU_SAMPLE_SECONDS = float(os.environ.get("TFX_BENCH_CPU_SECONDS", "10"))  # 1024 ch x 10 s slice, best of 3, + a 2x slice
SECONDARY_SECONDS = float(os.environ.get("TFX_BENCH_SECONDARY_SECONDS", "60"))  # signal length of configs 3-5
ALGO_BYTES_PER_SAMPLE = 8  # read x f32 once + write y f32 once


def sos_coefficients():
    import scipy.signal as sps

    return sps.butter(2 * SECTIONS, 5000.0 / (FS / 2), output="sos")  # == LoButterworth(5000, order=8), K = 4


def measured_hbm_peak() -> tuple[float, str]:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic():
    """dram bytes per launch of the main kernel from the committed ncu --set full capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler:
    """Samples SM clock / power / throttle reasons of one GPU WHILE the timed region runs.

    NVML from a Python thread every ~5 ms (the timed region can be as short as 0.1 s at 8 GPUs;
    `nvidia-smi -lms` is too coarse for that); falls back to an nvidia-smi subprocess when
    pynvml is unavailable.  CUDA calls made by the main thread release the GIL."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.samples: list[tuple[float, float, int]] = []
        self.sm_max = None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None
        self._handle = None
        self._smi = None
        try:
            import pynvml

            pynvml.nvmlInit()
            # NVML enumerates physical GPUs; honour CUDA_VISIBLE_DEVICES when it is a list of indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = gpu_index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if gpu_index < len(ids) and ids[gpu_index].isdigit():
                    phys = int(ids[gpu_index])
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM))
            self._nvml = pynvml
        except Exception:
            self._nvml = None

    def _loop(self):
        nv = self._nvml
        while not self._stop.is_set():
            try:
                clk = float(nv.nvmlDeviceGetClockInfo(self._handle, nv.NVML_CLOCK_SM))
                pw = nv.nvmlDeviceGetPowerUsage(self._handle) / 1000.0
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._handle))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._handle))
                self.samples.append((clk, pw, rs))
            except Exception:
                pass
            self._stop.wait(0.005)

    def start(self):
        if self._nvml is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
            return
        try:
            self._smi = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                 "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self._smi = None

    def stop(self) -> dict:
        if self._nvml is not None:
            self._stop.set()
            if self._thread is not None:
                self._thread.join(timeout=2)
            sm = [c for c, _, _ in self.samples]
            reasons = set()
            for _, _, rs in self.samples:
                for bit, name in self.REASONS.items():
                    if rs & bit:
                        reasons.add(name)
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.sm_max,
                    "power_w_max": max((p for _, p, _ in self.samples), default=None), "samples": len(sm),
                    "reasons": sorted(reasons), "source": "nvml, 5 ms period, timed region only"}
        if self._smi is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["clock sampling unavailable"]}
        self._smi.terminate()
        try:
            out = self._smi.communicate(timeout=5)[0]
        except Exception:
            self._smi.kill()
            out = ""
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            r = [c.strip() for c in line.split(",")]
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); smax.append(float(r[1])); power.append(float(r[2]))
            except ValueError:
                continue
            for name, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons),
                "source": "nvidia-smi -lms 20, timed region only"}


def pcie_probe(dev, world: int, nbytes: int = 1 << 30) -> dict:
    """Pinned-host <-> device copy rates of THIS rank while every rank copies at the same time: the ceiling of
    the e2e path at this GPU count (1 GiB each way; H2D alone, D2H alone, both at once on two streams)."""
    import torch
    import torch.distributed as dist

    n = nbytes // 4
    h_in = torch.empty(n, dtype=torch.float32, pin_memory=True)
    h_in.fill_(1.0)
    h_out = torch.empty(n, dtype=torch.float32, pin_memory=True)
    d_a = torch.empty(n, dtype=torch.float32, device=dev)
    d_b = torch.ones(n, dtype=torch.float32, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def h2d():
        with torch.cuda.stream(s1):
            d_a.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_b, non_blocking=True)

    def both():
        h2d()
        d2h()

    out = {}
    for name, fn in (("h2d", h2d), ("d2h", d2h), ("duplex_each_way", both)):
        fn()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            fn()
        torch.cuda.synchronize(dev)
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        out[name + "_GBps_per_gpu"] = round(2 * nbytes / float(dt.item()) / 1e9, 2)
    out["duplex_Gsamples_s_all_gpus"] = round(world * out["duplex_each_way_GBps_per_gpu"] / 4, 2)
    out["how"] = f"{nbytes >> 20} MiB pinned <-> device per direction and rank, all {world} rank(s) at once, slowest rank"
    del h_in, h_out, d_a, d_b
    return out


def secondary_configs(dev, world: int, rank: int, peak: float) -> dict:
    """BASELINE configs 3, 4 and 5 on this GPU count, channels sharded like the headline job; every
    entry carries its own HBM roofline.  For config 5 also the gathered form of SURVEY.md 8e: kernel only, kernel +
    chunk-overlapped NCCL all-gather of the output, and the link rate that implies.  An entry that fails reports the
    error instead of taking the headline line down with it."""
    import numpy as np
    import torch
    import torch.distributed as dist

    import torchfx_b200 as fx
    from oracle import oracle
    from torchfx_b200 import _native, _ops
    from torchfx_b200.dist import all_gather_channels, filter_and_gather, shard_bounds

    T = int(SECONDARY_SECONDS * FS)
    out: dict = {"seconds": SECONDARY_SECONDS, "n_gpus": world}

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _native.kernel_launches()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), (_native.kernel_launches() - l0) // reps

    def roof(bytes_per_gpu, ms):
        a = bytes_per_gpu / ms / 1e6
        return {"bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak, "per_gpu": True, "traffic": None}

    def rel(a, b):
        return float(np.abs(a - b).max() / np.abs(b).max())

    def guarded(name, fn):
        try:
            out[name] = fn()
        except Exception as e:  # keep the headline line alive
            out[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.empty_cache()

    # ---- config 4: fused pipe chain LoButterworth | ParametricEQ | HiShelving, 2048 channels ---------------
    def cfg4():
        C = 2048
        lo, hi = shard_bounds(C, world, rank)
        x = torch.empty((hi - lo, T), device=dev).normal_(0, 0.1)
        mk = lambda: [fx.filter.LoButterworth(5000, order=4, fs=FS), fx.filter.ParametricEQ(1000, q=2.0, gain=3.0, fs=FS),
                      fx.filter.HiShelving(8000, q=0.707, gain=2.0, gain_scale="db", fs=FS)]
        chain = mk()
        ms, nl = timed(lambda: (fx.Wave(x, FS, device=dev) | chain[0] | chain[1] | chain[2]).ys)
        n_chk = min(T, 1 << 17)
        y = (fx.Wave(x[:2, :n_chk], FS, device=dev) | chain[0] | chain[1] | chain[2]).ys
        sos = np.vstack([f._sos.numpy() for f in chain])
        want, _, _ = oracle.sos_cascade(x[:2, :n_chk].cpu().numpy(), sos)
        return {"workload": f"fused LoButterworth(5000,4) | ParametricEQ(1000,2,3) | HiShelving(8000,.707,2 dB), {C} ch x {SECONDARY_SECONDS:g} s, TFX_PREC_AUTO",
                "value": C * T / ms / 1e6, "unit": "Gsamples/s", "ms": ms, "launches_per_call": nl, "channels_per_gpu": hi - lo,
                "roofline": roof(8.0 * (hi - lo) * T, ms), "parity_rel_err": rel(y.cpu().numpy(), want)}

    # ---- config 5 (stack): LogFilterBank(32) over 256 channels -> [32, 256, T], sharded by channel ----------
    def cfg5_stack():
        C, N = 256, 32
        lo, hi = shard_bounds(C, world, rank)
        Cr = hi - lo
        x = torch.empty((Cr, T), device=dev).normal_(0, 0.1)
        bank = fx.filter.LogFilterBank(n_bands=N, f_min=20.0, f_max=20000.0, q=1.414, fs=FS)

        def run():
            bank.reset_state()
            return bank(x)

        ms, nl = timed(run)
        n_chk = min(T, 1 << 16)
        bank.reset_state()
        y = bank(x[:2, :n_chk])
        want = oracle.filterbank_stack(x[:2, :n_chk].cpu().numpy(), np.stack([f._sos.numpy() for f in bank.filters]))
        res = {"workload": f"LogFilterBank(32, 20 Hz .. 20 kHz) x {C} ch -> [32, {C}, T] (8192 lanes), {SECONDARY_SECONDS:g} s, TFX_PREC_AUTO",
               "value": N * C * T / ms / 1e6, "unit": "G lane-samples/s", "ms": ms, "launches_per_call": nl, "channels_per_gpu": Cr,
               "roofline": roof(4.0 * N * Cr * T * (1 + 1 / N), ms),
               "parity_rel_err": max(rel(y[b].cpu().numpy(), want[b]) for b in range(N))}
        del y
        if world > 1 and C % world == 0:
            # every rank ends up with the whole [32, 256, T].  Time chunks of 10 s; the gather of chunk i (one
            # all_gather of the rank's [32, C/P, n] block + one strided copy into band-major order) runs on a side stream
            # under the filtering of chunk i + 1.
            torch.cuda.empty_cache()  # the [32, 256, T] result is 94 GB: give the cached blocks of the timing runs back first
            full = torch.empty((N, C, T), dtype=torch.float32, device=dev)
            comm = torch.cuda.Stream(device=dev)
            chunk = max(FS, T // 6)  # 10 s chunks: long enough for the bank to split a chunk in time (the 20 Hz band forgets in ~0.65 s)
            stage = torch.empty((world, N, Cr, chunk), dtype=torch.float32, device=dev)  # one buffer: gather and copy-out are ordered on the comm stream

            def run_gather():
                bank.reset_state()
                cur = torch.cuda.current_stream(dev)
                for i, t0 in enumerate(range(0, T, chunk)):
                    n = min(chunk, T - t0)
                    yb = bank(x[:, t0:t0 + n])  # [N, Cr, n] contiguous, state carried by the bank's children
                    comm.wait_stream(cur)
                    with torch.cuda.stream(comm):
                        st = stage if n == chunk else torch.empty((world, N, Cr, n), dtype=torch.float32, device=dev)
                        dist.all_gather_into_tensor(st, yb)  # ONE collective per chunk: every rank's [N, Cr, n] block
                        # rank-major -> band-major: full[b, r * Cr + c, t0 + t] = st[r, b, c, t]
                        full[:, :, t0:t0 + n].unflatten(1, (world, Cr)).copy_(st.permute(1, 0, 2, 3))
                    yb.record_stream(comm)
                cur.wait_stream(comm)
                return full

            g_ms, _ = timed(run_gather, reps=2)
            recv = 4.0 * N * C * T * (world - 1) / world  # bytes every rank receives over NVLink
            n_cmp = min(T, chunk + chunk // 2)  # compare across the first chunk boundary (state carried) with one unchunked call
            bank.reset_state()
            ref = bank(x[:, :n_cmp])
            got = run_gather()[:, lo:hi, :n_cmp]
            res["gathered"] = {"kernel_only_ms": ms, "kernel_plus_overlapped_all_gather_ms": g_ms, "value": N * C * T / g_ms / 1e6,
                               "unit": "G lane-samples/s", "bytes_received_per_gpu": recv, "nvlink_GBps_per_gpu": recv / g_ms / 1e6,
                               "gather_over_kernel": g_ms / ms, "chunk_samples": chunk,
                               "rel_diff_vs_unchunked": float((got - ref).abs().max() / ref.abs().max())}
        return res

    # ---- config 5 (sum): parallel `+` biquads -----------------------------------------------------------------
    def cfg5_sum(nb, C, label):
        def go():
            lo, hi = shard_bounds(C, world, rank)
            x = torch.empty((hi - lo, T), device=dev).normal_(0, 0.1)
            fl = ([fx.filter.BiquadBPF(200.0 * 1.7 ** i, 1.414, FS) for i in range(nb)] if nb == 8 else
                  [fx.filter.BiquadBPF(20.0 * (1000.0 ** (i / (nb - 1.0))), 1.414, FS) for i in range(nb)])
            comb = fx.filter._base.ParallelFilterCombination(*fl)

            def run():
                for f in fl:
                    f.reset_state()
                return comb(x)

            ms, nl = timed(run)
            n_chk = min(T, 1 << 16)
            for f in fl:
                f.reset_state()
            y = comb(x[:2, :n_chk])
            want = oracle.filterbank_sum(x[:2, :n_chk].cpu().numpy(), np.stack([f._sos.numpy() for f in fl]))
            return {"workload": f"{label}: {nb} BiquadBPF added with `+` over {C} ch ({nb * C} biquad lanes), {SECONDARY_SECONDS:g} s, TFX_PREC_AUTO",
                    "value": C * T / ms / 1e6, "unit": "Gsamples/s", "biquad_lane_value": nb * C * T / ms / 1e6, "ms": ms,
                    "launches_per_call": nl, "channels_per_gpu": hi - lo, "roofline": roof(8.0 * (hi - lo) * T, ms),
                    "fma_per_channel_sample": 5 * nb, "parity_rel_err": rel(y.cpu().numpy(), want)}
        return go

    # ---- config 3: 256-channel overlap-save FIR, 65 536-tap reverb IR --------------------------------------
    def cfg3():
        C, K = 256, 65536
        lo, hi = shard_bounds(C, world, rank)
        rng = np.random.default_rng(7)
        ir = rng.standard_normal(K) * np.exp(-np.arange(K) / 8000.0)
        ir = (ir / np.sqrt((ir ** 2).sum())).astype(np.float32)
        x = torch.empty((hi - lo, T), device=dev).normal_(0, 0.1)
        f = fx.filter.FIR(ir)
        ms, nl = timed(lambda: f(x))
        n_chk = min(T, 150000)
        y = f(x[:3, :n_chk])
        want = oracle.fir_causal(x[:3, :n_chk].cpu().numpy(), ir)
        r = roof(8.0 * (hi - lo) * T, ms)
        flops = 2.0 * 130 * (hi - lo) * T  # ~130 fp32 instructions per sample, counted as FMAs
        return {"workload": f"FIR overlap-save, 65 536-tap decaying-noise IR, {C} ch x {SECONDARY_SECONDS:g} s",
                "value": C * T / ms / 1e6, "unit": "Gsamples/s", "ms": ms, "launches_per_call": nl, "channels_per_gpu": hi - lo,
                "roofline": r, "fp32_TFLOPs_per_gpu_est": flops / ms / 1e9, "note": "fp32-issue bound, not HBM bound (DRAM traffic 10.9 B/sample, profiles/r2_fir.md)",
                "parity_rel_err": rel(y.cpu().numpy(), want)}

    guarded("cfg4_fused_chain", cfg4)
    guarded("cfg5_filterbank_stack", cfg5_stack)
    guarded("cfg5_sum_8x1024", cfg5_sum(8, 1024, "config 5 secondary"))
    guarded("cfg5_sum_32x256", cfg5_sum(32, 256, "config 5 read literally"))
    guarded("cfg3_fir_65536", cfg3)
    return out


def reference_cpu_step(ext, x32, sos_t):
    """What the reference does for one filter call on a CPU tensor: up-cast to float64
    (_ops.py:142), fused DF1 cascade in the native extension (iir_cpu.cpp:64-159), cast back
    to the input dtype (filter/iir.py:176).  Zero initial state (fresh filter)."""
    import torch

    C = x32.shape[0]
    K = sos_t.shape[0]
    sx = torch.zeros(K, C, 2, dtype=torch.float64)
    sy = torch.zeros(K, C, 2, dtype=torch.float64)
    y, _, _ = ext.sos_forward(x32.to(torch.float64), sos_t, sos_t, sx, sy)
    return y.to(torch.float32)


def cpu_reference_runner():
    """Returns (kind, cores, fn(x32)->y32): the unmodified reference extension when
    oracle/_ref is present, else the oracle port."""
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    sos = sos_coefficients()
    sos_t = torch.from_numpy(sos).contiguous()
    from oracle import ref_loader

    if ref_loader.have_ref_ext():
        try:
            ext = ref_loader.load_ref_ext()
            return "reference", cores, (lambda x32: reference_cpu_step(ext, x32, sos_t))
        except Exception as e:  # pragma: no cover
            print(f"[bench] reference extension failed to load ({e}); using the oracle port", file=sys.stderr)
    from oracle import oracle

    oracle.build()
    return "port", cores, (lambda x32: torch.from_numpy(oracle.sos_cascade(x32.numpy(), sos)[0]))


def time_cpu_baseline(sample_seconds: float, reps: int = 3, warm: bool = True):
    """the reference CPU path on a 1024 ch x 10 s slice, best of 3 after a warm-up, plus ONE run of a
    2x longer slice to confirm that the time scales linearly (so the slice extrapolates to the 10-minute job)."""
    import torch

    kind, cores, fn = cpu_reference_runner()
    T = int(sample_seconds * FS)
    g = torch.Generator().manual_seed(1234)
    x = 0.1 * torch.randn(CHANNELS, 2 * T, generator=g)
    if warm:
        fn(x[:, : min(T, 4800)])
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        fn(x[:, :T])
        best = min(best, time.perf_counter() - t0)
    t0 = time.perf_counter()
    fn(x)
    t2 = time.perf_counter() - t0
    return {
        "value": CHANNELS * T / best / 1e9,
        "unit": "Gsamples/s",
        "cores": cores,
        "kind": kind,
        "sample": f"{CHANNELS} ch x {sample_seconds:g} s @48kHz f32 ({CHANNELS * T / 1e6:.1f} Msamples), f32->f64->f32 casts included, best of {reps}",
        "seconds": best,
        "linearity": {"slice_2x_seconds": t2, "slice_2x_value": CHANNELS * 2 * T / t2 / 1e9, "time_ratio_2x_over_1x": t2 / best},
    }


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    T = int(CPU_SAMPLE_SECONDS * FS)
    import torch

    kind, cores, fn = cpu_reference_runner()
    g = torch.Generator().manual_seed(1234)
    x = 0.1 * torch.randn(CHANNELS, T, generator=g)
    for _ in range(max(args.warmup, 1)):
        fn(x[:, : max(T // 8, 1)])
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        fn(x)
        times.append(time.perf_counter() - t0)
    total = sum(times)
    value = CHANNELS * T * args.steps / total / 1e9
    sample = f"each step = {CHANNELS} ch x {CPU_SAMPLE_SECONDS:g} s slice of the workload ({CHANNELS * T / 1e6:.1f} Msamples), reference CPU kernel on {cores} threads"
    line = {
        "impl": "reference",
        "metric": "Gsamples/sec, 1024-ch 4-SOS cascade @48kHz",
        "value": value,
        "unit": "Gsamples/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args.gpus),  # key-for-key the dict of the product arm: same workload, same sharding plan
        "arm_note": "the CPU kernel is timed on a bounded slice of this workload (cpu_baseline.sample); rank 0 only",
        "cpu_baseline": {"value": value, "unit": "Gsamples/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus: int) -> dict:
    """The workload both arms are quoted on (identical dict in the product line and the reference line)."""
    cfg = {
        "workload": "BASELINE configs[1]: 1024-ch x 10-min @48kHz float32, 4-section SOS biquad cascade (LoButterworth(5000, order=8))",
        "channels": CHANNELS,
        "samples_per_channel": int(SECONDS * FS),
        "sections": SECTIONS,
        "fs": FS,
    }
    cfg.update({
        "channels_per_gpu": CHANNELS // n_gpus,
        "sharding": f"channels over {n_gpus} rank(s), no data-path collective",
        "in_place": True,
        "l2": "inputs larger than L2 (no flush needed): %.1f GB per GPU" % (CHANNELS // n_gpus * SECONDS * FS * 4 / 1e9),
        "precision": "TFX_PREC_AUTO",
    })
    if SECONDS != 600:
        cfg["reduced"] = f"TFX_BENCH_SECONDS={SECONDS:g} (not the BASELINE size)"
    return cfg


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip BASELINE configs 3-5 (the `secondary` object)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import torchfx_b200 as fx  # noqa: F401
    from torchfx_b200 import _native, _ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from torchfx_b200.dist import shard_bounds

    lo, hi = shard_bounds(CHANNELS, world, rank)
    C = hi - lo
    T = int(SECONDS * FS)
    sos_np = sos_coefficients()
    sos = torch.from_numpy(sos_np).contiguous()

    # ---- synthetic input, resident in HBM before the timed region ------------------------
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.empty((C, T), dtype=torch.float32, device=dev)
    x.normal_(0.0, 0.1, generator=gen)

    # ---- parity spot-check of the very kernel configuration that is timed ----------------
    from oracle import oracle

    pick = [0, C // 3, C - 1]
    n_chk = min(T, 1 << 18)
    x_chk = x[pick, :n_chk].cpu().numpy()

    def step():
        _ops.sos_cascade_(x, sos, None, None, out=x)  # in place: 118 GB in + 118 GB out does not fit 180 GB

    step()  # warm-up step 1 doubles as the parity check
    torch.cuda.synchronize()
    want, _, _ = oracle.sos_cascade(x_chk, sos_np)
    got = x[pick, :n_chk].cpu().numpy()
    parity = float(np.abs(got - want).max() / np.abs(want).max())
    if not parity < 1e-5:
        raise SystemExit(f"parity failure before timing: rel-to-max error {parity}")
    x.normal_(0.0, 0.1, generator=gen)
    for _ in range(args.warmup - 1):
        step()
    torch.cuda.synchronize()

    # ---- timed region ---------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    launches0 = _native.kernel_launches()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    launches = _native.kernel_launches() - launches0
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    nl = torch.tensor([float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(nl, op=dist.ReduceOp.SUM)
    total_ms = float(ms.item())
    ms_per_step = total_ms / args.steps
    value = CHANNELS * T / (ms_per_step * 1e-3) / 1e9

    # roofline of the dominant kernel on this rank (per GPU): algorithmic bytes / step time.
    # The step is warm-up launch + main launch; the warm-up touches < 0.1 % of the samples,
    # so the step time is charged entirely to the main kernel (conservative).
    peak, peak_src = measured_hbm_peak()
    achieved = ALGO_BYTES_PER_SAMPLE * C * T / (ms_per_step * 1e-3) / 1e9
    traffic = recorded_traffic()
    traffic_bytes = None
    if traffic and traffic.get("traffic_over_algorithmic"):
        traffic_bytes = traffic["traffic_over_algorithmic"] * ALGO_BYTES_PER_SAMPLE * C * T
    roofline = {
        "bound": "hbm",
        "achieved": achieved,
        "peak": peak,
        "unit": "GB/s",
        "frac": achieved / peak,
        "traffic": traffic_bytes,
        "traffic_source": (traffic or {}).get("capture"),
        "kernel": "sos_tile_kernel<float,float,4> (32-channel x 64-sample cp.async tiles, persistent warps)",
        "algorithmic_bytes_per_launch": ALGO_BYTES_PER_SAMPLE * C * T,
        "peak_source": peak_src,
        "per_gpu": True,
    }
    del x
    torch.cuda.empty_cache()

    # ---- end to end through the host-buffer C-ABI entry (pinned host -> GPU -> pinned host) --
    e2e = None
    if not args.no_e2e:
        from torchfx_b200.dist import bind_to_gpu_numa_node

        # pinned buffers and the copy-driving thread local to the GPU's PCIe root (round 1: e2e did not scale past ~16)
        numa = {"skipped": "TFX_BENCH_NO_NUMA"} if os.environ.get("TFX_BENCH_NO_NUMA") else bind_to_gpu_numa_node(local_rank)
        probe = pcie_probe(dev, world)
        Te = int(E2E_SECONDS * FS)
        xh = torch.empty((C, Te), dtype=torch.float32, pin_memory=True)
        xh.normal_(0.0, 0.1)
        yh = torch.empty((C, Te), dtype=torch.float32, pin_memory=True)
        lib = _native.load()

        def e2e_step():
            _native.check(lib.tfx_sos_cascade_host_f32(xh.data_ptr(), yh.data_ptr(), C, Te, Te, Te, sos.data_ptr(), SECTIONS,
                                                        None, None, 0, 0, local_rank))

        e2e_step()
        chk = oracle.sos_cascade(xh[:2, : 1 << 16].numpy(), sos_np)[0]
        e2e_par = float(np.abs(yh[:2, : 1 << 16].numpy() - chk).max() / np.abs(chk).max())
        if not e2e_par < 1e-5:
            raise SystemExit(f"e2e parity failure: {e2e_par}")
        e2e_steps = 3
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()  # synchronous: returns when y_host is complete
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {
            "value": CHANNELS * Te * e2e_steps / float(dt.item()) / 1e9,
            "unit": "Gsamples/s",
            "h2d_bytes_per_step": 4 * CHANNELS * Te,
            "d2h_bytes_per_step": 4 * CHANNELS * Te,
            "api": "tfx_sos_cascade_host_f32 (pinned host buffers, chunked H2D/kernel/D2H overlap)",
            "sample": f"{CHANNELS} ch x {E2E_SECONDS:g} s slice per step ({e2e_steps} steps), PCIe-bound",
            "numa_binding_rank0": numa,
            "pcie_probe": probe,
            "of_pcie_duplex_ceiling": None,
        }
        if probe.get("duplex_Gsamples_s_all_gpus"):
            e2e["of_pcie_duplex_ceiling"] = round(e2e["value"] / probe["duplex_Gsamples_s_all_gpus"], 3)
        del xh, yh

    secondary = None
    if not args.no_secondary:
        secondary = secondary_configs(dev, world, rank, peak)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu = time_cpu_baseline(CPU_SAMPLE_SECONDS)

    line = {
        "metric": "Gsamples/sec, 1024-ch 4-SOS cascade @48kHz",
        "value": value,
        "unit": "Gsamples/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms_per_step,
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(world),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "e2e": e2e,
        "gpu_launches": int(nl.item()),
        "parity_rel_err": parity,
        "secondary": secondary,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
