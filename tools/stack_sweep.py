"""Developer tool: A/B the libraries in build/variants on the STACK filterbank (config 5 shape)."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
child = r'''
import os, sys, json, torch
sys.path.insert(0, %r)
import torchfx_b200 as fx
from torchfx_b200 import _ops
def t(fn, reps=4):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
out = {}
for C, T, N in ((256, 2880000, 32), (1024, 1440000, 8), (64, 2880000, 32)):
    x = torch.empty((C, T), dtype=torch.float32, device="cuda").normal_(0, 0.1, generator=torch.Generator(device="cuda").manual_seed(3))
    bank = fx.filter.LogFilterBank(n_bands=N, f_min=20.0, f_max=20000.0, q=1.414, fs=48000)
    for prec in ("f32", "auto"):
        _ops.set_default_precision(prec)
        ms = t(lambda: (bank.reset_state(), bank(x))[1])
        out[f"{C}x{T}x{N}_{prec}"] = [round(ms, 3), round(4 * N * C * T * (1 + 1 / N) / ms / 1e6)]
    bank.reset_state()
    y = bank(x)
    out[f"{C}x{T}x{N}_chk"] = sum(float(y[b].abs().sum(dtype=torch.float64)) for b in range(N))
    del y, x
    torch.cuda.empty_cache()
print(json.dumps(out))
''' % ROOT
libs = sorted(glob.glob(os.path.join(ROOT, "build", "variants", "lib_*.so")))
res = {}
for lib in libs:
    env = dict(os.environ, TFX_B200_LIB=lib)
    p = subprocess.run([sys.executable, "-c", child], env=env, capture_output=True, text=True, timeout=900)
    name = os.path.basename(lib)
    try:
        res[name] = json.loads(p.stdout.strip().splitlines()[-1])
    except Exception:
        res[name] = {"error": (p.stderr or p.stdout)[-600:]}
    print(name, res[name], flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "stack_variants.json"), "w"), indent=1)
