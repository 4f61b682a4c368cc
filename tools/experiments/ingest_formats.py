"""Format-step kernel alone: device time and HBM throughput per source format (device-resident
raw buffers, separate pass forced).  Prints one JSON line."""
import ctypes as C
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
os.environ["RSB_PCM_UNFUSED"] = "1"
from resampler_b200 import Attenuation, FirBatch, Latency, PcmFormat, _lib  # noqa: E402
from resampler_b200.fir import MEM_DEVICE  # noqa: E402

lib = _lib.load()
n, frames, ch = 512, 44100 * 20, 2
out = {}
for fmt in (PcmFormat.F32, PcmFormat.S32, PcmFormat.S24, PcmFormat.S16, PcmFormat.U8):
    for src_ch in (2, 1):
        bps = fmt.bytes_per_sample()
        raw_bytes = frames * src_ch * bps
        stride = (raw_bytes + 15) & ~15
        d_raw = lib.rsb_alloc_device(0, stride * n)
        b = FirBatch(n, ch, 44100, 48000, Latency.Sample64, Attenuation.Db90)
        cap = ((int(frames / b.ratio()) + 8) * ch + 3) & ~3
        d_out = lib.rsb_alloc_device(0, cap * 4 * n)
        assert d_raw and d_out
        one = np.random.default_rng(1).integers(0, 256, raw_bytes, dtype=np.uint8)
        if fmt == PcmFormat.F32:
            one = np.random.default_rng(1).uniform(-1, 1, frames * src_ch).astype(np.float32).view(np.uint8)
        lib.rsb_memcpy(0, d_raw, one.ctypes.data, raw_bytes, 0)
        for s in range(1, n):
            lib.rsb_memcpy(0, d_raw + s * stride, d_raw, raw_bytes, 2)
        ms = []
        for _ in range(4):
            b.reset(-1)
            b.process_pcm_ptrs([d_raw + s * stride for s in range(n)], [frames] * n, fmt, src_ch, 512, 0,
                               [d_out + s * cap * 4 for s in range(n)], [cap] * n, memspace=MEM_DEVICE)
            ms.append(b.last_ingest_ms())
        t = float(np.mean(ms[1:]))
        nbytes = n * (raw_bytes + frames * ch * 4)
        out[f"{fmt.name}_{src_ch}ch"] = {"ms": round(t, 4), "GB/s": round(nbytes / t / 1e6, 1),
                                         "frac_of_6554": round(nbytes / t / 1e6 / 6554.2, 3)}
        lib.rsb_free_device(0, d_raw)
        lib.rsb_free_device(0, d_out)
        b.close()
print(json.dumps(out))
