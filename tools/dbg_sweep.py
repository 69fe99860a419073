"""Developer tool: time one shape under different TFX_DEBUG / flag settings (subprocess each)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
child = r'''
import os, sys, json, torch, scipy.signal as sps
sys.path.insert(0, %r)
from torchfx_b200 import _ops
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
out = {}
for case in os.environ["VS_CASES"].split(";"):
    C, T, K, prec, mode = case.split(","); C, T, K = int(C), int(T), int(K)
    x = torch.empty((C, T), dtype=torch.float32, device="cuda").normal_(0, 0.1)
    y = x if "ip" in mode else torch.empty_like(x)
    sos = torch.from_numpy(sps.butter(2 * K, 5000 / 24000, output="sos")).contiguous()
    ms = t(lambda: _ops.sos_cascade_(x, sos, None, None, out=y, precision=prec, no_tma=("gen" in mode), force_tma=("gen" not in mode)))
    out[case] = round(8 * C * T / ms / 1e6, 1)
    del x, y
print(json.dumps(out))
''' % ROOT
cases = os.environ.get("VS_CASES", "1024,2880000,1,f32,ip;1024,2880000,1,f32,oop;1024,2880128,1,f32,ip;1024,2883584,1,f32,ip;1024,2880000,1,f32,gen-ip;1024,2880000,1,f32,gen-oop")
for dbg in os.environ.get("DBGS", "0,1,2,3").split(","):
    env = dict(os.environ, TFX_DEBUG=dbg, VS_CASES=cases)
    p = subprocess.run([sys.executable, "-c", child], env=env, capture_output=True, text=True, timeout=600)
    print("TFX_DEBUG=" + dbg, p.stdout.strip().splitlines()[-1] if p.stdout.strip() else p.stderr[-300:], flush=True)
