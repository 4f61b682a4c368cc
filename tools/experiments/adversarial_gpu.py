"""GPU check of the fp32-ordering difference for an adversarial full-scale input: +-1 samples
whose sign pattern follows the filter's centre-phase row (tiled), so that aligned windows reach
|y| ~ sum|g| = 2.65.  Tensor / FFMA2 kernels against the bit-exact kernel."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from resampler_b200 import Attenuation, FirBatch, Kernel, Latency  # noqa: E402

n, ch, frames, taps = 64, 2, 30000, 128
b0 = FirBatch(n, ch, 44100, 48000, Latency.Sample64, Attenuation.Db90, kernel=Kernel.EXACT)
table = b0.coeffs().reshape(1024, taps)
rng = np.random.default_rng(3)
xs = []
for s in range(n):
    row = np.sign(table[(s * 16) % 1024]).astype(np.float32)
    row[row == 0] = 1.0
    x = np.tile(row, frames // taps + 1)[:frames]
    xs.append(np.repeat(x, ch))            # the same signal in both channels
ref = b0.process(xs, 512 * ch)
b0.close()
peak = max(float(np.abs(o).max()) for o in ref["out"])
for kern in (Kernel.TENSOR, Kernel.FAST):
    b = FirBatch(n, ch, 44100, 48000, Latency.Sample64, Attenuation.Db90, kernel=kern)
    r = b.process(xs, 512 * ch)
    worst = max(float(np.abs(a.astype(np.float64) - e).max()) for a, e in zip(r["out"], ref["out"]))
    above = sum(int((np.abs(a.astype(np.float64) - e) > 1e-6).sum()) for a, e in zip(r["out"], ref["out"]))
    total = sum(a.size for a in r["out"])
    print(f"{b.last_kernel().name}: peak |y| {peak:.3f}, max |diff to EXACT| {worst:.3e}, above 1e-6: {above} of {total}")
    b.close()
