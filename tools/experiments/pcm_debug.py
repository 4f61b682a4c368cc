"""Debug aid: per-batch error of device-resident PCM batches against the oracle."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib as O
from resampler_b200 import Attenuation, FirBatch, Kernel, Latency, PcmFormat, _lib
from resampler_b200.fir import MEM_DEVICE
from test_gpu_pcm import raw_samples

def run(kernel, fmt, src_ch, seed):
    ch, n, frames = 2, 64, 6007
    lib = _lib.load()
    rng = np.random.default_rng(seed)
    raws = [raw_samples(rng, fmt, frames * src_ch) for _ in range(n)]
    raws2 = [raw_samples(rng, fmt, 1500 * src_ch) for _ in range(n)]
    bps = fmt.bytes_per_sample()
    stride = (frames * src_ch * bps + 15) & ~15
    d_raw = lib.rsb_alloc_device(0, stride * n + 16)
    batch = FirBatch(n, ch, 44100, 48000, Latency.Sample64, Attenuation.Db90, kernel=kernel)
    cap = int(frames * ch / batch.ratio()) + 4 * batch.buffer_size_output(); cap -= cap % 4
    d_out = lib.rsb_alloc_device(0, cap * 4 * n)
    res = []
    fs = [O.OracleFir(ch, 44100, 48000, 3, 1) for _ in range(n)]
    for rs, nf in ((raws, frames), (raws2, 1500)):
        for i, r in enumerate(rs):
            b = np.ascontiguousarray(r).view(np.uint8)
            lib.rsb_memcpy(0, d_raw + i * stride, b.ctypes.data, b.nbytes, 0)
        cons, prod, calls = batch.process_pcm_ptrs([d_raw + i * stride for i in range(n)], [nf] * n, fmt, src_ch,
                                                   512, 0, [d_out + i * cap * 4 for i in range(n)], [cap] * n,
                                                   memspace=MEM_DEVICE)
        errs = []
        for i in range(n):
            ref = fs[i].process(O.pcm_to_f32(rs[i], int(fmt), 1 if src_ch == ch else ch), 512)
            got = np.empty(prod[i], np.float32)
            lib.rsb_memcpy(0, got.ctypes.data, d_out + i * cap * 4, got.nbytes, 1)
            d = np.abs(got.astype(np.float64) - ref["out"])
            errs.append((float(d.max()), int(np.argmax(d)), int((d > 1e-6).sum())))
        bad = [(i, e) for i, e in enumerate(errs) if e[0] > 1e-6]
        res.append((max(e[0] for e in errs), len(bad), bad[:4]))
    lib.rsb_free_device(0, d_raw); lib.rsb_free_device(0, d_out); batch.close()
    return res

for kernel in (Kernel.TENSOR, Kernel.EXACT, Kernel.FAST):
    for fmt, sc in ((PcmFormat.S24, 1), (PcmFormat.S16, 1), (PcmFormat.S24, 2)):
        for seed in (79, 80):
            print(kernel.name, fmt.name, sc, seed, run(kernel, fmt, sc, seed), flush=True)
