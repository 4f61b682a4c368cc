"""CPU simulation: how far can fp32 summation-ORDER differences go for adversarial input?
x = +-1 with the sign pattern of the filter row (|y| reaches sum|g| ~ 2.65, the largest output a
[-1, 1] signal can produce), against the oracle's AVX-512 order.  Models: (a) the reference's own
scalar order (fir/mod.rs:47-62), (b) the tensor kernel's order (3xTF32, small terms first, hi*hi
outside-in, products exact, sequential fp32 accumulation).  No GPU needed."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "tools" / "experiments"))
import oracle_lib as O  # noqa: E402
from lo_fp16_sim import acc_f32, tf32  # noqa: E402

taps = 128
table = O.design_table(44100, 48000, 3, 1)
rng = np.random.default_rng(1)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
p1 = rng.integers(0, 1023, n)
fr = rng.uniform(0, 1, n).astype(np.float32)
for name in ("adversarial sign pattern", "full-scale noise"):
    g = (table[p1 + 1] * fr[:, None] + table[p1] * (np.float32(1) - fr)[:, None]).astype(np.float32)
    x = np.sign(g).astype(np.float32) if name.startswith("adv") else rng.uniform(-1, 1, (n, taps)).astype(np.float32)
    avx = np.array([O.convolve(x[i], table[p1[i]], table[p1[i] + 1], fr[i], O.CONV_AVX512) for i in range(n)])
    sca = np.array([O.convolve(x[i], table[p1[i]], table[p1[i] + 1], fr[i], O.CONV_SCALAR) for i in range(n)])
    xh, gh = tf32(x), tf32(g)
    xl, gl = tf32(x - xh), tf32(g - gh)
    order = []
    a, b = 0, taps - 1
    while a <= b:
        order.append(a)
        if b != a:
            order.append(b)
        a += 1
        b -= 1
    small = np.empty((n, 2 * taps), np.float64)
    small[:, 0::2] = xl.astype(np.float64) * gh
    small[:, 1::2] = xh.astype(np.float64) * gl
    big = (xh.astype(np.float64) * gh)[:, np.array(order)]
    tc = acc_f32(np.concatenate([small, big], axis=1))
    print(f"{name}: max |y| {np.abs(avx).max():.3f}; vs the oracle's AVX-512 order: "
          f"reference scalar order {np.abs(sca.astype(np.float64) - avx).max():.3e}, "
          f"tensor-kernel order {np.abs(tc.astype(np.float64) - avx).max():.3e}")
