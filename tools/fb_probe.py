import sys, torch, ctypes
sys.path.insert(0, '/root/repo')
import torchfx_b200 as fx
from torchfx_b200 import _ops, _native
dev = torch.device("cuda:0")
C, T, N = 256, 2880000, 32
x = torch.empty((C, T), device=dev).normal_(0, 0.1)
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
bank = fx.filter.LogFilterBank(n_bands=N, f_min=20.0, f_max=20000.0, q=1.414, fs=48000)
bank.compute_coefficients()
lib = _native.load(); err = ctypes.c_double()
precs = [lib.tfx_sos_auto_precision(f._sos.data_ptr(), 1, ctypes.byref(err)) for f in bank.filters]
print("auto prec per band:", precs)
for prec in ("f32", "f64", "auto"):
    _ops.set_default_precision(prec)
    ms = t(lambda: (bank.reset_state(), bank(x))[1])
    print(prec, round(ms, 2), "ms", round(N*C*T/ms/1e6, 1), "G lane-samples/s")
# N=8 sum
_ops.set_default_precision("f32")
x2 = torch.empty((1024, T), device=dev).normal_(0, 0.1)
fl = [fx.filter.BiquadBPF(200.0 * 1.7 ** i, 1.414, 48000) for i in range(8)]
comb = fx.filter._base.ParallelFilterCombination(*fl)
for prec in ("f32", "f64"):
    _ops.set_default_precision(prec)
    def run():
        for f in fl: f.reset_state()
        return comb(x2)
    ms = t(run)
    print("sum", prec, round(ms, 2), "ms", round(1024*T/ms/1e6, 1), "Gsamples/s")
