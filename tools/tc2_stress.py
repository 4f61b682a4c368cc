"""Stress test of the tensor kernel's synchronisation: many repetitions of configurations whose
tiles are short / long / partial, every run compared bit for bit with the first one (the kernel
is deterministic) and with EXACT within 1e-6.  A hang shows up as the caller's timeout.
Under gpurun:  timeout 600 python tools/tc2_stress.py [reps]"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from resampler_b200 import Attenuation, FirBatch, Kernel, Latency  # noqa: E402

CASES = [
    # ch, in_hz, out_hz, latency, call frames, streams, frames
    (1, 16000, 48000, 1, 160, 512, 64000),     # short tiles (5 K steps), many items per CTA
    (2, 44100, 48000, 3, 512, 1024, 30000),    # headline geometry, 16 groups
    (2, 48000, 44100, 3, 512, 300, 20000),     # down-sampling, partial last group
    (8, 96000, 48000, 2, 512, 100, 12000),     # 8 channels
    (1, 8000, 48000, 0, 64, 700, 9000),        # 16 taps, ratio 1/6: tiles barely advance
    (2, 44100, 48000, 3, 512, 65, 3000),       # tiny: fewer items than CTAs
]


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    for case in CASES:
        ch, in_hz, out_hz, lat, call, n, frames = case
        rng = np.random.default_rng(7)
        xs = [rng.uniform(-1, 1, frames * ch).astype(np.float32) for _ in range(n)]
        e = FirBatch(n, ch, in_hz, out_hz, Latency(lat), Attenuation.Db90, kernel=Kernel.EXACT)
        want = np.stack([np.array(o, copy=True) for o in e.process(xs, call * ch, 0)["out"]])
        e.close()
        b = FirBatch(n, ch, in_hz, out_hz, Latency(lat), Attenuation.Db90, kernel=Kernel.TENSOR)
        first = None
        t0 = time.time()
        for r in range(reps):
            b.reset(-1)
            got = np.stack(b.process(xs, call * ch, 0)["out"])
            assert b.last_kernel() == Kernel.TENSOR, b.last_kernel()
            if first is None:
                first = got.copy()
                worst = float(np.abs(got.astype(np.float64) - want).max())
                assert worst <= 1e-6, (case, worst)
            else:
                assert np.array_equal(got.view(np.uint32), first.view(np.uint32)), (case, r)
        b.close()
        print(f"ok {case} x{reps}: max |diff| vs EXACT {worst:.3e}, {time.time() - t0:.1f} s", flush=True)


if __name__ == "__main__":
    main()
