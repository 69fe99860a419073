"""Thread-level numpy model of the 8192-point transform of csrc/fir_ols8k.cu (256 threads, two CTAs per SM).

Same idea as tools/fft16k_model.py: executes the per-thread program (register slots, shared-memory slots, twiddle
exponents) on the CPU, checks it against numpy.fft and counts bank conflicts of every 64-bit shared-memory instruction.

Forward (decimation in frequency), n = n1*512 + d2*32 + d3*2 + jl:
  P1  radix-16 over n1  CTA-wide  thread t <-> positions r = 2t, 2t+1 (pair A, B)      -> sub-transform k1 (512 points)
  P2  radix-16 over d2  warp      warp w owns sub-transforms 2w, 2w+1; lane = (hs = lane >> 4, sp = lane & 15)
  P3  radix-16 over d3  warp      lane = (hs, k2); position j = jl (the halves of the pair)
  P4  radix-2 over jl ACROSS the halves of a pair (scalar adds), in registers
Output bin k1 + 16*(k2 + 16*(k3 + 16*r2)) is half r2 of the pair at index slot*256 + tid of the row.
"""
import numpy as np
from fft16k_model import radix16_dif, radix16_dit_inv, SLOT_K, Conflicts

N = 8192


def xslot8(k2, u4):
    return k2 * 16 + ((u4 ^ k2) & 15)


def forward(x, cf=None):
    cf = cf or Conflicts()
    sm = np.zeros((16, 256, 2), dtype=complex)  # [sub-transform][pair slot][half]
    tid = np.arange(256)
    v = np.empty((16, 256, 2), dtype=complex)
    for m in range(16):
        for hf in range(2):
            v[m, :, hf] = x[512 * m + 2 * tid + hf]
    v = radix16_dif(v, np.stack([2 * tid, 2 * tid + 1], axis=1), 512)
    for s in range(16):
        sm[SLOT_K[s], tid, :] = v[s]
        for w in range(8):
            cf.check("P1 store", list(range(32 * w, 32 * w + 32)))
    lane = np.arange(32)
    hs, l16 = lane >> 4, lane & 15
    row = np.zeros((4096, 2), dtype=complex)
    binmap = np.zeros((4096, 2), dtype=int)
    for w in range(8):
        k1 = 2 * w + hs  # per lane
        R = [sm[2 * w], sm[2 * w + 1]]
        # region base in pair slots: sub-transform index * 512 (256 re + 256 im) -> bank-neutral
        v = np.stack([np.where((hs == 0)[:, None], R[0][m * 16 + l16], R[1][m * 16 + l16]) for m in range(16)])
        for m in range(16):
            cf.check("P2 load", (2 * w + hs) * 512 + m * 16 + l16)
        v = radix16_dif(v, np.stack([2 * l16, 2 * l16 + 1], axis=1), 32)
        R2 = [np.zeros((256, 2), dtype=complex), np.zeros((256, 2), dtype=complex)]
        for s in range(16):
            k2 = SLOT_K[s]
            for h in range(2):
                R2[h][xslot8(k2, l16[hs == h])] = v[s][hs == h]
            cf.check("P2 store", (2 * w + hs) * 512 + xslot8(k2, l16))
        # P3: lane = (hs, k2 = l16)
        k2 = l16
        v = np.stack([np.where((hs == 0)[:, None], R2[0][xslot8(k2, m)], R2[1][xslot8(k2, m)]) for m in range(16)])
        for m in range(16):
            cf.check("P3 load", (2 * w + hs) * 512 + xslot8(k2, m))
        j = np.stack([np.zeros(32, dtype=int), np.ones(32, dtype=int)], axis=1)
        v = radix16_dif(v, j, 2)
        for s in range(16):
            k3 = SLOT_K[s]
            A, B = v[s][:, 0], v[s][:, 1]
            o0, o1 = A + B, A - B
            pidx = s * 256 + (w * 32 + lane)
            row[pidx, 0], row[pidx, 1] = o0, o1
            for r2 in range(2):
                binmap[pidx, r2] = k1 + 16 * (k2 + 16 * (k3 + 16 * r2))
    return row, binmap, cf


def inverse(row):
    lane = np.arange(32)
    hs, l16 = lane >> 4, lane & 15
    tid = np.arange(256)
    sm = np.zeros((16, 256, 2), dtype=complex)
    for w in range(8):
        k2 = l16
        v = np.empty((16, 32, 2), dtype=complex)
        for s in range(16):
            pidx = s * 256 + (w * 32 + lane)
            o0, o1 = row[pidx, 0], row[pidx, 1]
            v[s, :, 0], v[s, :, 1] = o0 + o1, o0 - o1
        j = np.stack([np.zeros(32, dtype=int), np.ones(32, dtype=int)], axis=1)
        v = radix16_dit_inv(v, j, 2)
        R2 = [np.zeros((256, 2), dtype=complex), np.zeros((256, 2), dtype=complex)]
        for m in range(16):
            for h in range(2):
                R2[h][xslot8(k2[hs == h], m)] = v[m][hs == h]
        v = np.stack([np.where((hs == 0)[:, None], R2[0][xslot8(SLOT_K[s], l16)], R2[1][xslot8(SLOT_K[s], l16)]) for s in range(16)])
        v = radix16_dit_inv(v, np.stack([2 * l16, 2 * l16 + 1], axis=1), 32)
        for m in range(16):
            for h in range(2):
                sm[2 * w + h, m * 16 + l16[hs == h]] = v[m][hs == h]
    v = np.stack([sm[SLOT_K[s], tid, :] for s in range(16)])
    v = radix16_dit_inv(v, np.stack([2 * tid, 2 * tid + 1], axis=1), 512)
    x = np.zeros(N, dtype=complex)
    for m in range(16):
        for hf in range(2):
            x[512 * m + 2 * tid + hf] = v[m, :, hf]
    return x


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    row, binmap, cf = forward(x)
    X = np.fft.fft(x)
    assert sorted(binmap.reshape(-1).tolist()) == list(range(N)), "bin map is not a permutation"
    err = np.abs(row - X[binmap]).max() / np.abs(X).max()
    print("forward  max rel err vs numpy.fft:", err)
    back = inverse(row) / N
    print("inverse  max abs err:", np.abs(back - x).max())
    print("worst bank conflict (ways) per 64-bit shared-memory instruction:", cf.worst)
    assert err < 1e-12 and np.abs(back - x).max() < 1e-12
