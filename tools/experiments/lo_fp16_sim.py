"""CPU simulation for a next-round option of the tensor kernel: keep the
x_lo ring as fp16 (two frames per 32-bit TMEM column, kind::f16, K = 16) instead of TF32.
Question: does x_lo(fp16) * g_hi(fp16), accumulated in fp32 next to x_hi*g_lo and x_hi*g_hi
(TF32), stay inside the 1e-6 bar?  Model: products exact, fp32 sequential accumulation in the
kernel's order (small terms first, then hi*hi outside-in).  No GPU needed.

    python tools/experiments/lo_fp16_sim.py [windows]
"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from resampler_b200 import _lib  # noqa: E402


def tf32(a):
    """cvt.rna.tf32.f32: round to nearest (ties away) keeping 10 explicit mantissa bits."""
    u = a.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x1000) & 0xFFFFE000
    return u.astype(np.uint32).view(np.float32)


def acc_f32(terms):
    """sequential fp32 accumulation along axis 1"""
    s = np.zeros(terms.shape[0], np.float32)
    for j in range(terms.shape[1]):
        s = (s + terms[:, j].astype(np.float32)).astype(np.float32)
    return s


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    taps = 128
    lib = _lib.load()
    table = np.empty(1024 * taps, np.float32)
    bits = C.c_uint32(0)
    assert lib.rsb_host_design_table(44100, 48000, 3, 1, table.ctypes.data_as(_lib.f32p), table.size,
                                     C.byref(bits)) == 0
    table = table.reshape(1024, taps)
    rng = np.random.default_rng(0)
    for name, scale in (("full-scale noise", 1.0), ("|x| <= 0.01", 0.01)):
        x = (rng.uniform(-1, 1, (n, taps)) * scale).astype(np.float32)
        p1 = rng.integers(0, 1023, n)
        fr = rng.uniform(0, 1, n).astype(np.float32)
        g = (table[p1 + 1] * fr[:, None] + table[p1] * (np.float32(1) - fr)[:, None]).astype(np.float32)
        ref = (x.astype(np.float64) * g.astype(np.float64)).sum(axis=1)
        xh, gh = tf32(x), tf32(g)
        xl, gl = tf32(x - xh), tf32(g - gh)
        order = []                      # hi*hi outside-in: front, back, front + 1, back - 1, ...
        a, b = 0, taps - 1
        while a <= b:
            order.append(a)
            if b != a:
                order.append(b)
            a += 1
            b -= 1
        order = np.array(order)

        def total(lo_term):
            small = np.empty((n, 2 * taps), np.float64)
            small[:, 0::2] = lo_term
            small[:, 1::2] = xh.astype(np.float64) * gl.astype(np.float64)
            big = (xh.astype(np.float64) * gh.astype(np.float64))[:, order]
            return acc_f32(np.concatenate([small, big], axis=1))

        cur = total(xl.astype(np.float64) * gh.astype(np.float64))
        xl16 = (x - xh).astype(np.float16).astype(np.float64)
        gh16 = gh.astype(np.float16).astype(np.float64)
        var = total(xl16 * gh16)
        print(f"{name}: {n} windows of {taps} taps")
        print(f"  3xTF32 (today)            max |err| {np.abs(cur - ref).max():.3e}")
        print(f"  x_lo, g_hi as fp16        max |err| {np.abs(var - ref).max():.3e}   "
              f"max |variant - today| {np.abs(var.astype(np.float64) - cur).max():.3e}")


if __name__ == "__main__":
    main()
