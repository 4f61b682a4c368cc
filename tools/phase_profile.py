"""Per-phase cycle breakdown of the fast kernel (debug counters).  Under gpurun:
   python tools/phase_profile.py [seconds]"""
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from resampler_b200 import Attenuation, FirBatch, Kernel, Latency, _lib  # noqa: E402
from resampler_b200.fir import FLAG_ASYNC, MEM_DEVICE, DeviceBuffer  # noqa: E402

# two-CTA kernel: setup, tma_issue+zero, row_build, tma_wait, deinterleave, product, store
# warp-specialised kernel: producer wait_empty / issue+meta / wait_tma / deinterleave+handover,
#                          consumer group 0 wait_full / product+store, group 1 the same
NAMES = ["P.wait_empty", "P.issue+meta", "P.wait_tma", "P.deint+handover", "C0.wait_full",
         "C0.product+store", "C1.wait_full", "C1.product+store"]


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 6.0
    ch = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    n, frames = 2048 // ch, int(44100 * seconds)
    lib = _lib.load()
    b = FirBatch(n, ch, 44100, 48000, Latency.Sample64, Attenuation.Db90, kernel=Kernel.FAST)
    d_in = DeviceBuffer(0, n * frames * ch)
    lib.rsb_fill_synthetic(0, d_in.ptr, 0, n, frames, ch, 44100, 0x5EED)
    out_stride = (int(frames * 48000 / 44100) + 8) * ch
    d_out = DeviceBuffer(0, n * out_stride)
    args = ([d_in.ptr + 4 * s * frames * ch for s in range(n)], [frames * ch] * n, 512 * ch, 0,
            [d_out.ptr + 4 * s * out_stride for s in range(n)], [out_stride] * n)
    for _ in range(2):
        b.reset(-1)
        b.process_ptrs(*args, memspace=MEM_DEVICE, flags=FLAG_ASYNC)
    b.sync()
    out = (C.c_uint64 * 8)()
    lib.rsb_debug_phase_cycles(b._h, 1, out)
    b.reset(-1)
    b.process_ptrs(*args, memspace=MEM_DEVICE, flags=FLAG_ASYNC)
    b.sync()
    conv_ms = float(b.conv_times_ms(1)[0])
    lib.rsb_debug_phase_cycles(b._h, 0, out)
    cyc = np.array(out[:], dtype=np.float64)
    tot = cyc.sum()
    kcyc = conv_ms * 1e-3 * 1.965e9 * 147     # kernel cycles x CTAs
    res = {"conv_ms": conv_ms,
           "pct_of_kernel_time": {NAMES[i]: round(100 * cyc[i] / kcyc, 1) for i in range(8)}}
    print(json.dumps(res))
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "phase_profile.json").write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
