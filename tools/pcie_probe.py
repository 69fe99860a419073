"""Developer tool: the PCIe ceiling of the host-streaming path.  Times pinned H2D, D2H and both
at once (two streams), as 1-D copies and as the 2-D (row-pitched time-chunk) copies the streaming
driver issues, so that the e2e number can be read against what the link can do."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
C, T, CH = 1024, 60 * 48000, 65536
host_in = torch.empty((C, T), dtype=torch.float32).pin_memory()
host_out = torch.empty((C, T), dtype=torch.float32).pin_memory()
host_in.normal_()
dev = torch.empty((C, T), dtype=torch.float32, device="cuda")
dev2 = torch.empty((C, T), dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
nbytes = 4 * C * T
res = {}
def h2d_1d():
    with torch.cuda.stream(s1): dev.copy_(host_in, non_blocking=True)
def d2h_1d():
    with torch.cuda.stream(s2): host_out.copy_(dev2, non_blocking=True)
def both_1d():
    h2d_1d(); d2h_1d()
res["h2d_1d_GBps"] = round(nbytes / timed(h2d_1d) / 1e9, 2)
res["d2h_1d_GBps"] = round(nbytes / timed(d2h_1d) / 1e9, 2)
res["duplex_1d_GBps_each_way"] = round(nbytes / timed(both_1d) / 1e9, 2)
nchunk = (T + CH - 1) // CH
def h2d_2d():
    with torch.cuda.stream(s1):
        for i in range(nchunk): dev[:, i * CH:(i + 1) * CH].copy_(host_in[:, i * CH:(i + 1) * CH], non_blocking=True)
def d2h_2d():
    with torch.cuda.stream(s2):
        for i in range(nchunk): host_out[:, i * CH:(i + 1) * CH].copy_(dev2[:, i * CH:(i + 1) * CH], non_blocking=True)
def both_2d():
    h2d_2d(); d2h_2d()
res["h2d_2d_GBps"] = round(nbytes / timed(h2d_2d) / 1e9, 2)
res["d2h_2d_GBps"] = round(nbytes / timed(d2h_2d) / 1e9, 2)
res["duplex_2d_GBps_each_way"] = round(nbytes / timed(both_2d) / 1e9, 2)
res["duplex_2d_Gsamples_s"] = round(res["duplex_2d_GBps_each_way"] / 4, 2)
res["duplex_1d_Gsamples_s"] = round(res["duplex_1d_GBps_each_way"] / 4, 2)
print(json.dumps({"pcie_probe": res, "shape": [C, T], "chunk_T": CH}))
