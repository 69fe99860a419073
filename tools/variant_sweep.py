"""Developer tool: A/B the kernel-geometry variants in build/variants on the GPU box."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = os.environ.get("VS_CASES", "1024,2880000,4,f32;1024,2880000,1,f32;128,23040000,4,f32;1024,2880000,4,f64;1024,2880000,8,f32;1024,28800000,4,f32")
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
    parts = case.split(","); C, T, K, prec = int(parts[0]), int(parts[1]), int(parts[2]), parts[3]; tma = len(parts) > 4 and parts[4] == "tma"; pk = len(parts) > 4 and parts[4] == "pk"; nt = len(parts) > 4 and parts[4] == "notile"
    x = torch.empty((C, T), dtype=torch.float32, device="cuda").normal_(0, 0.1)
    sos = torch.from_numpy(sps.butter(2 * K, 5000 / 24000, output="sos")).contiguous()
    ms = t(lambda: _ops.sos_cascade_(x, sos, None, None, out=x, precision=prec, force_tma=tma, packed=pk, no_tile=nt))
    out[case] = round(8 * C * T / ms / 1e6, 1)
    del x
print(json.dumps(out))
''' % ROOT
res = {}
libs = sorted(glob.glob(os.path.join(ROOT, "build", "variants", "lib_*.so")))
for lib in libs:
    env = dict(os.environ, TFX_B200_LIB=lib, VS_CASES=CASES)
    p = subprocess.run([sys.executable, "-c", child], env=env, capture_output=True, text=True, timeout=600)
    name = os.path.basename(lib)
    try:
        res[name] = json.loads(p.stdout.strip().splitlines()[-1])
    except Exception:
        res[name] = {"error": (p.stderr or p.stdout)[-400:]}
    print(name, res[name], flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "variants.json"), "w"), indent=1)
