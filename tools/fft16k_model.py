"""Thread-level numpy model of the 16384-point transform used by csrc/fir_ols16k.cu.

Developer tool (not product, not oracle): it executes the SAME per-thread program the CUDA kernel runs --
which thread holds which elements in which register slot, every shared-memory address of every exchange,
every twiddle exponent -- so that the index maps can be checked on the CPU (against numpy.fft) and so that
the bank behaviour of every shared-memory instruction can be counted before GPU time is spent.

Decomposition (forward, decimation in frequency), n = n1*1024 + d2*64 + d3*4 + jh*2 + jl:
  P1  radix-16 over n1   CTA-wide   thread t <-> positions r = 2t, 2t+1 (packed pair A, B)     -> sub-FFT k1 (= warp)
  P2  radix-16 over d2   warp       lane l  <-> pair-slot sp = l  (s = d3*4 + jh*2 + {0,1})
  P3  radix-16 over d3   warp       lane l  <-> (jh = l >> 4, k2 = l & 15)
  P4  radix-2 over jh (packed) + radix-2 over jl (across the halves of a pair, scalar)
                                    lane l  <-> (h = l >> 4, k2 = l & 15), k3 = 8h + i
Output bin f = k1 + 16*(k2 + 16*(k3 + 16*(r + 2*r2))); stored as bin PAIRS (r2 = 0, 1) at pair index
(i*2 + r)*512 + tid of the row.
"""
import numpy as np

N = 16384


def W(n, e):
    return np.exp(-2j * np.pi * (np.asarray(e) % n) / n)


def radix16_dif(v, j, q):
    """v: [16, ...] complex; j: position inside the 16q-point group (array ok).  Slot 4r+r2 <- output k = r + 4 r2."""
    out = np.empty_like(v)
    for m in range(4):
        u = [v[m + 4 * s] for s in range(4)]
        for r in range(4):
            o = sum(u[s] * W(4, s * r) for s in range(4))
            out[m + 4 * r] = o * W(16 * q, (j + m * q) * r)
    v = out.copy()
    for a in range(4):
        u = [v[4 * a + mm] for mm in range(4)]
        for r2 in range(4):
            out[4 * a + r2] = sum(u[mm] * W(4, mm * r2) for mm in range(4)) * W(4 * q, j * r2)
    return out


def radix16_dit_inv(v, j, q):
    """exact inverse of radix16_dif up to the factor 16"""
    out = np.empty_like(v)
    for a in range(4):
        u = [v[4 * a + r2] * np.conj(W(4 * q, j * r2)) for r2 in range(4)]
        for mm in range(4):
            out[4 * a + mm] = sum(u[r2] * np.conj(W(4, mm * r2)) for r2 in range(4))
    v = out.copy()
    for m in range(4):
        u = [v[m + 4 * r] * np.conj(W(16 * q, (j + m * q) * r)) for r in range(4)]
        for s in range(4):
            out[m + 4 * s] = sum(u[r] * np.conj(W(4, s * r)) for r in range(4))
    return out


SLOT_K = [(s >> 2) + 4 * (s & 3) for s in range(16)]  # register slot -> output digit


def xslot(k2, u5):
    """pair-slot of (k2, u5) inside a warp's 512-slot region for the P2->P3 and P3->P4 exchanges"""
    return k2 * 32 + ((u5 & 16) | ((u5 ^ k2) & 15))


class Conflicts:
    def __init__(self):
        self.worst = {}

    def check(self, name, slots):
        """slots: [32] pair-slot (8-byte word) index per lane of ONE 64-bit shared-memory instruction"""
        for half in range(2):
            s = np.asarray(slots[16 * half: 16 * half + 16]) % 16
            ways = np.bincount(s, minlength=16).max()
            self.worst[name] = max(self.worst.get(name, 1), int(ways))


def forward(x, cf=None):
    """x: [16384] complex (re = channel a, im = channel b).  Returns row[8192, 2] (bin pairs) and the bin map."""
    cf = cf or Conflicts()
    sm = np.zeros((16, 512, 2), dtype=complex)  # [warp region][pair-slot][half]
    tid = np.arange(512)
    # ---- P1: CTA-wide, thread t <-> r = 2t + {0,1}
    v = np.empty((16, 512, 2), dtype=complex)
    for m in range(16):
        for hf in range(2):
            v[m, :, hf] = x[1024 * m + 2 * tid + hf]
    j = np.stack([2 * tid, 2 * tid + 1], axis=1)
    v = radix16_dif(v, j, 1024)
    for s in range(16):
        sm[SLOT_K[s], tid, :] = v[s]  # region k1, slot u = t
        for w in range(16):
            cf.check("P1 store", list(range(32 * w, 32 * w + 32)))
    # ---- warp-local
    lane = np.arange(32)
    row = np.zeros((8192, 2), dtype=complex)
    binmap = np.zeros((8192, 2), dtype=int)
    for w in range(16):
        k1 = w
        R = sm[w]
        # P2: lane <-> sp = lane ; element (d2 = m, sp)
        v = np.stack([R[m * 32 + lane] for m in range(16)])  # [16, 32, 2]
        for m in range(16):
            cf.check("P2 load", m * 32 + lane)
        j = np.stack([2 * lane, 2 * lane + 1], axis=1)  # s = 2 sp + half
        v = radix16_dif(v, j, 64)
        R2 = np.zeros_like(R)
        for s in range(16):
            k2 = SLOT_K[s]
            R2[xslot(k2, lane)] = v[s]
            cf.check("P2 store", xslot(k2, lane))
        # P3: lane <-> (jh = lane >> 4, k2 = lane & 15); element (k2, d3 = m, jh) at u5 = 2 m + jh
        jh, k2 = lane >> 4, lane & 15
        v = np.stack([R2[xslot(k2, 2 * m + jh)] for m in range(16)])
        for m in range(16):
            cf.check("P3 load", xslot(k2, 2 * m + jh))
        j = np.stack([2 * jh, 2 * jh + 1], axis=1)
        v = radix16_dif(v, j, 4)
        R3 = np.zeros_like(R)
        for s in range(16):
            k3 = SLOT_K[s]
            R3[xslot(k2, 2 * k3 + jh)] = v[s]
            cf.check("P3 store", xslot(k2, 2 * k3 + jh))
        # P4: lane <-> (h = lane >> 4, k2 = lane & 15); k3 = 8h + i; registers (i, jh)
        h, k2 = lane >> 4, lane & 15
        for i in range(8):
            k3 = 8 * h + i
            v0 = R3[xslot(k2, 2 * k3 + 0)]  # [32, 2]
            v1 = R3[xslot(k2, 2 * k3 + 1)]
            cf.check("P4 load", xslot(k2, 2 * k3 + 0))
            cf.check("P4 load", xslot(k2, 2 * k3 + 1))
            for r, pk in ((0, v0 + v1), (1, v0 - v1)):  # packed radix-2 over jh
                A, B = pk[:, 0], pk[:, 1]
                if r == 1:
                    B = B * (-1j)  # W_4^{jl * r}
                o0, o1 = A + B, A - B  # scalar radix-2 over jl: r2 = 0, 1
                pidx = (i * 2 + r) * 512 + (w * 32 + lane)
                row[pidx, 0], row[pidx, 1] = o0, o1
                for r2 in range(2):
                    binmap[pidx, r2] = k1 + 16 * (k2 + 16 * (k3 + 16 * (r + 2 * r2)))
    return row, binmap, cf


def inverse(row):
    """mirror image of forward(); returns the UNSCALED inverse transform (x * N)"""
    lane = np.arange(32)
    tid = np.arange(512)
    sm = np.zeros((16, 512, 2), dtype=complex)
    for w in range(16):
        h, k2 = lane >> 4, lane & 15
        R3 = np.zeros((512, 2), dtype=complex)
        for i in range(8):
            k3 = 8 * h + i
            pk = []
            for r in range(2):
                pidx = (i * 2 + r) * 512 + (w * 32 + lane)
                o0, o1 = row[pidx, 0], row[pidx, 1]
                A, B = o0 + o1, o0 - o1
                if r == 1:
                    B = B * (1j)
                pk.append(np.stack([A, B], axis=1))
            v0, v1 = pk[0] + pk[1], pk[0] - pk[1]
            R3[xslot(k2, 2 * k3 + 0)] = v0
            R3[xslot(k2, 2 * k3 + 1)] = v1
        jh, k2 = lane >> 4, lane & 15
        v = np.stack([R3[xslot(k2, 2 * SLOT_K[s] + jh)] for s in range(16)])
        j = np.stack([2 * jh, 2 * jh + 1], axis=1)
        v = radix16_dit_inv(v, j, 4)
        R2 = np.zeros((512, 2), dtype=complex)
        for m in range(16):
            R2[xslot(k2, 2 * m + jh)] = v[m]
        v = np.stack([R2[xslot(SLOT_K[s], lane)] for s in range(16)])
        j = np.stack([2 * lane, 2 * lane + 1], axis=1)
        v = radix16_dit_inv(v, j, 64)
        for m in range(16):
            sm[w, m * 32 + lane] = v[m]
    v = np.stack([sm[SLOT_K[s], tid, :] for s in range(16)])
    j = np.stack([2 * tid, 2 * tid + 1], axis=1)
    v = radix16_dit_inv(v, j, 1024)
    x = np.zeros(N, dtype=complex)
    for m in range(16):
        for hf in range(2):
            x[1024 * m + 2 * tid + hf] = v[m, :, hf]
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
