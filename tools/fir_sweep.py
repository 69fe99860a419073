"""Developer tool: A/B the libraries in build/variants on the FIR overlap-save path (config 3 and shorter IRs)."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
child = r'''
import os, sys, json, numpy as np, torch
sys.path.insert(0, %r)
import torchfx_b200 as fx
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
out = {}
x = torch.empty((256, 2880000), dtype=torch.float32, device="cuda").normal_(0, 0.1, generator=torch.Generator(device="cuda").manual_seed(3))
rng = np.random.default_rng(7)
for K in (65536, 20000, 4096, 1024):
    ir = rng.standard_normal(K) * np.exp(-np.arange(K) / (K / 8.0))
    ir = (ir / np.sqrt((ir ** 2).sum())).astype(np.float32)
    f = fx.filter.FIR(ir)
    ms = t(lambda: f(x))
    y = f(x)
    out[str(K)] = [round(ms, 3), round(x.numel() / ms / 1e6, 1), float(y.abs().sum(dtype=torch.float64))]
    del y
print(json.dumps(out))
''' % ROOT
for lib in sorted(glob.glob(os.path.join(ROOT, "build", "variants", "lib_*.so"))):
    env = dict(os.environ, TFX_B200_LIB=lib)
    p = subprocess.run([sys.executable, "-c", child], env=env, capture_output=True, text=True, timeout=900)
    try:
        print(os.path.basename(lib), json.loads(p.stdout.strip().splitlines()[-1]), flush=True)
    except Exception:
        print(os.path.basename(lib), "error", (p.stderr or p.stdout)[-600:], flush=True)
