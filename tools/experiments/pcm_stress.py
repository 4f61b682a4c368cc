"""Stress: fused raw-s16 tensor path against the bit-exact kernel, short tiles / many items per CTA
(the geometry that exposed the tensor kernel's two synchronisation bugs), many seeds."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from resampler_b200 import Attenuation, FirBatch, Kernel, Latency, PcmFormat  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
bad = 0
for ch, src_ch, in_hz, out_hz, lat, n in ((1, 1, 16000, 48000, 1, 1024), (2, 2, 44100, 48000, 3, 256),
                                          (2, 1, 48000, 44100, 3, 192)):
    worst = 0.0
    for seed in range(reps):
        rng = np.random.default_rng(seed)
        frames = 2000 + 37 * seed
        raws = [rng.integers(-32768, 32768, frames * src_ch, dtype=np.int16) for _ in range(n)]
        res = {}
        for kern in (Kernel.TENSOR, Kernel.EXACT):
            b = FirBatch(n, ch, in_hz, out_hz, Latency(lat), Attenuation.Db90, kernel=kern)
            r = b.process_pcm(raws, PcmFormat.S16, src_ch, call_len=160 * ch)
            assert b.last_pcm_fused() == (kern == Kernel.TENSOR)
            res[kern] = r
            b.close()
        for s in range(n):
            a, e = res[Kernel.TENSOR]["out"][s], res[Kernel.EXACT]["out"][s]
            assert a.size == e.size
            d = float(np.max(np.abs(a.astype(np.float64) - e))) if a.size else 0.0
            worst = max(worst, d)
            if d > 1e-6:
                bad += 1
                print("MISMATCH", ch, src_ch, seed, s, d, flush=True)
    print(f"ch {ch} src {src_ch} {in_hz}->{out_hz}: {reps} seeds x {n} streams, worst {worst:.3e}", flush=True)
print("bad streams:", bad)
