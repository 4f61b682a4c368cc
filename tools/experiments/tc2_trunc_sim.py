"""CPU model of the tensor kernel's accumulation: fp16 hi/lo operands (exact products), per-MMA
fp32 accumulate with TRUNCATION (round toward zero) or round-to-nearest, K steps of 16 frames on
the tile's chunk grid, in the kernel's issue order.  Compares orders / tile widths against the
f64 value of the reference formula and against the AVX-512-order f32 result."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from resampler_b200 import _lib  # noqa: E402

lib = _lib.load()
IN, OUT, LAT, TAPS = 44100, 48000, 3, 128


def table():
    buf = np.zeros(1024 * TAPS, np.float32)
    bits = C.c_uint32(0)
    assert lib.rsb_host_design_table(IN, OUT, LAT, 1, buf.ctypes.data_as(C.POINTER(C.c_float)), buf.size,
                                     C.byref(bits)) == 0
    return buf.reshape(1024, TAPS)


def rz32(a):
    """float64 -> float32 toward zero"""
    f = a.astype(np.float32)
    over = np.abs(f.astype(np.float64)) > np.abs(a)
    f[over] = np.nextafter(f[over], np.float32(0))
    return f


def split16(a, scale):
    A = (a.astype(np.float64) * scale)
    hi = A.astype(np.float16).astype(np.float64)
    lo = (A - hi).astype(np.float16).astype(np.float64)
    return hi, lo


def main():
    n_out, ntile = 64 * 400, 64
    T = table()
    k = np.arange(n_out)
    pos = k * (IN / OUT)
    v = np.floor(pos).astype(np.int64)
    ph = (pos - v) * 1024
    p1 = np.minimum(np.floor(ph), 1023).astype(np.int64)
    fr = (ph - p1).astype(np.float32)
    p2 = np.minimum(p1 + 1, 1023)
    g = (T[p2] * fr[:, None] + T[p1] * (np.float32(1) - fr)[:, None]).astype(np.float32)   # blended rows
    rng = np.random.default_rng(1)
    rows = 24
    x = rng.uniform(-1, 1, (rows, v[-1] + TAPS + 64)).astype(np.float32)
    ghi, glo = split16(g, 8192.0)
    res = {}
    mma_count = {}
    true = np.zeros((rows, n_out))
    for mode in ("rz64", "rn64", "rz_quarter", "rz64_inorder"):
        res[mode] = np.zeros((rows, n_out))
    xhi, xlo = split16(x, 16.0)
    for t0 in range(0, n_out, ntile):
        o = np.arange(t0, t0 + ntile)
        k0 = v[t0] - (v[t0] % 16)
        kt = ((v[t0 + ntile - 1] + TAPS - k0 + 15) // 16) * 16
        nks = kt // 16
        # dense banded G for the tile: [64][kt]
        Ghi = np.zeros((ntile, kt)); Glo = np.zeros((ntile, kt)); Gf = np.zeros((ntile, kt))
        for i, oo in enumerate(o):
            d = v[oo] - k0
            Ghi[i, d:d + TAPS] = ghi[oo]; Glo[i, d:d + TAPS] = glo[oo]; Gf[i, d:d + TAPS] = g[oo].astype(np.float64)
        Xh = xhi[:, k0:k0 + kt]; Xl = xlo[:, k0:k0 + kt]; Xf = x[:, k0:k0 + kt].astype(np.float64)
        true[:, o] = Xf @ Gf.T
        def steps_outside_in(n):
            f, b, order = 0, n - 1, []
            while f <= b:
                order.append(f)
                if b != f: order.append(b)
                f += 1; b -= 1
            return order
        oi = steps_outside_in(nks)
        def partial(A, B, s):
            return A[:, 16 * s:16 * s + 16] @ B[:, 16 * s:16 * s + 16].T
        for mode, rnd in (("rz64", rz32), ("rn64", lambda a: a.astype(np.float32))):
            D = np.zeros((rows, ntile), np.float32)
            for s in range(nks):
                D = rnd(D.astype(np.float64) + partial(Xl, Ghi, s))
                D = rnd(D.astype(np.float64) + partial(Xh, Glo, s))
            for s in oi:
                D = rnd(D.astype(np.float64) + partial(Xh, Ghi, s))
            res[mode][:, o] = D.astype(np.float64) / (16.0 * 8192.0)
        # in-order main pass
        D = np.zeros((rows, ntile), np.float32)
        for s in range(nks):
            D = rz32(D.astype(np.float64) + partial(Xl, Ghi, s))
            D = rz32(D.astype(np.float64) + partial(Xh, Glo, s))
        for s in range(nks):
            D = rz32(D.astype(np.float64) + partial(Xh, Ghi, s))
        res["rz64_inorder"][:, o] = D.astype(np.float64) / (16.0 * 8192.0)
        # per-quarter order: each 16-output quarter adds its own central steps last
        D = np.zeros((rows, ntile), np.float32)
        for s in range(nks):
            D = rz32(D.astype(np.float64) + partial(Xl, Ghi, s))
            D = rz32(D.astype(np.float64) + partial(Xh, Glo, s))
        for q in range(4):
            cols = slice(16 * q, 16 * q + 16)
            cent = sorted({(v[t0 + 16 * q + i] - k0 + 64) // 16 for i in (0, 15)} |
                          {(v[t0 + 16 * q + i] - k0 + 63) // 16 for i in (0, 15)})
            rest = [s for s in oi if s not in cent]
            Dq = D[:, cols].copy()
            for s in rest + cent:
                Dq = rz32(Dq.astype(np.float64) + Xh[:, 16 * s:16 * s + 16] @ Ghi[cols, 16 * s:16 * s + 16].T)
            res["rz_quarter"][:, cols.start + t0:cols.stop + t0] = Dq.astype(np.float64) / (16.0 * 8192.0)

        # range-split order: every K step is issued "early" for the column quarters it is not
        # central for and "late" for those it is central for (MMAs over contiguous quarter ranges)
        for mode, nq in (("rz_split4", 4), ("rz_split2", 2)):
            w = ntile // nq
            D = np.zeros((rows, ntile), np.float32)
            for s in range(nks):
                D = rz32(D.astype(np.float64) + partial(Xl, Ghi, s))
                D = rz32(D.astype(np.float64) + partial(Xh, Glo, s))
            cent = []
            for q in range(nq):
                lo_c = (v[t0 + w * q] - k0 + 63) // 16
                hi_c = (v[t0 + w * q + w - 1] - k0 + 64) // 16
                cent.append(set(range(lo_c, hi_c + 1)))
            n_mma = 0
            for late in (False, True):
                for s in (oi if not late else sorted(set().union(*cent))):
                    qs = [q for q in range(nq) if (s in cent[q]) == late]
                    if not qs: continue
                    # contiguous ranges
                    ranges, start = [], qs[0]
                    for a, b in zip(qs, qs[1:] + [None]):
                        if b != a + 1:
                            ranges.append((start, a)); start = b
                    for (a, b) in ranges:
                        cols = slice(w * a, w * (b + 1))
                        D[:, cols] = rz32(D[:, cols].astype(np.float64) + Xh[:, 16 * s:16 * s + 16] @ Ghi[cols, 16 * s:16 * s + 16].T)
                        n_mma += 1
            res.setdefault(mode, np.zeros((rows, n_out)))[:, o] = D.astype(np.float64) / (16.0 * 8192.0)
            mma_count[mode] = mma_count.get(mode, 0) + n_mma
    print({k: v / (n_out // ntile) for k, v in mma_count.items()}, 'main-pass MMAs per tile (13 unsplit)')
    big = np.abs(true) > 0.25
    for mode, y in res.items():
        d = y - true
        sg = d * np.sign(true)
        print(f"{mode:14s} max {np.abs(d).max():.3e} rms {np.sqrt((d*d).mean()):.3e} "
              f"rel bias {(sg[big]/np.abs(true[big])).mean()*2**24:+.3f} x 2^-24")
        c = (sg[big] / np.abs(true[big])).mean()
        d2 = y * (1 - c) - true
        print(f"{'':14s} after bias compensation: max {np.abs(d2).max():.3e} rms {np.sqrt((d2*d2).mean()):.3e}")


if __name__ == "__main__":
    main()
