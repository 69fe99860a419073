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
CPU_SAMPLE_SECONDS = float(os.environ.get("TFX_BENCH_CPU_SECONDS", "5"))
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


def time_cpu_baseline(sample_seconds: float, reps: int = 1, warm: bool = True):
    import torch

    kind, cores, fn = cpu_reference_runner()
    T = int(sample_seconds * FS)
    g = torch.Generator().manual_seed(1234)
    x = 0.1 * torch.randn(CHANNELS, T, generator=g)
    if warm:
        fn(x[:, : min(T, 4800)])
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        fn(x)
        best = min(best, time.perf_counter() - t0)
    return {
        "value": CHANNELS * T / best / 1e9,
        "unit": "Gsamples/s",
        "cores": cores,
        "kind": kind,
        "sample": f"{CHANNELS} ch x {sample_seconds:g} s @48kHz f32 ({CHANNELS * T / 1e6:.1f} Msamples), f32->f64->f32 casts included, best of {reps}",
        "seconds": best,
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
        "config": workload_config(1, reference=True),
        "cpu_baseline": {"value": value, "unit": "Gsamples/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus: int, reference: bool = False) -> dict:
    cfg = {
        "workload": "BASELINE configs[1]: 1024-ch x 10-min @48kHz float32, 4-section SOS biquad cascade (LoButterworth(5000, order=8))",
        "channels": CHANNELS,
        "samples_per_channel": int(SECONDS * FS),
        "sections": SECTIONS,
        "fs": FS,
    }
    if reference:
        cfg["note"] = "reference arm: CPU kernel timed on a bounded slice of this workload (see cpu_baseline.sample)"
        return cfg
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
        }
        del xh, yh

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
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
