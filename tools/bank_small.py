"""Developer tool: small STACK / SUM filterbank calls that take the time-parallel warm-up kernel (target of compute-sanitizer
memcheck / racecheck runs); prints the difference to the serial warm-up."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchfx_b200 as fx
from torchfx_b200 import _native
from torchfx_b200.filter._sosbank import SosBank

freqs = [20.0 * (1000.0 ** (i / 31.0)) for i in range(32)]
for mode, C, T in (("stack", 32, 200000), ("sum", 5, 150001)):
    x = 0.1 * torch.randn(C, T, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    out = {}
    for env in ("TFX_BS_SERIAL_WARM", "TFX_BS_PARALLEL_WARM"):
        os.environ.pop("TFX_BS_SERIAL_WARM", None); os.environ.pop("TFX_BS_PARALLEL_WARM", None)
        os.environ[env] = "1"
        bank = SosBank([fx.filter.BiquadBPF(f, 1.414, 48000) for f in freqs], mode=mode)
        bank.flags = _native.TFX_FORCE_TILE
        n0 = _native.kernel_launches()
        out[env] = torch.cat([bank(x[:, :60032]), bank(x[:, 60032:])], dim=-1)
        launches = _native.kernel_launches() - n0
    d = float((out["TFX_BS_SERIAL_WARM"] - out["TFX_BS_PARALLEL_WARM"]).abs().max() / out["TFX_BS_SERIAL_WARM"].abs().max())
    print(mode, C, T, "launches", launches, "serial vs parallel warm-up:", d)
