"""Per-role cycle breakdown of the tensor kernel (debug counters of CTA 0).  Under gpurun:
   python tools/tc_profile.py [seconds] [channels]"""
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from resampler_b200 import Attenuation, FirBatch, Kernel, Latency, _lib  # noqa: E402
from resampler_b200.fir import FLAG_ASYNC, MEM_DEVICE, DeviceBuffer  # noqa: E402

NAMES = ["M.wait_x_full", "M.wait_g_full", "M.wait_d_empty", "M.issue", "M.commit+meta",
         "S.wait_x_empty", "S.wait_xs_full", "S.load+split", "E.wait_t_done", "E.ld+stage", "E.tma_store",
         "S.wait_st+arrive", "S.st", "kernel", "tiles", "S.other", "J.wait_t_done", "J.release",
         "G.wait_t_done", "G.issue", "E.wait_store_read", "E.stage_writes", "E.fence", "P.wait_item",
         "P.item_setup", "P.wait_xs_empty", "P.tma_issue", "x", "x", "x", "x", "x"]


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 6.0
    ch = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    n, frames = 2048 // ch, int(44100 * seconds)
    lib = _lib.load()
    b = FirBatch(n, ch, 44100, 48000, Latency.Sample64, Attenuation.Db90, kernel=Kernel.TENSOR)
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
    assert b.last_kernel() == Kernel.TENSOR
    out = (C.c_uint64 * 32)()
    lib.rsb_debug_tc_cycles(b._h, 1, out, 32)
    b.reset(-1)
    b.process_ptrs(*args, memspace=MEM_DEVICE, flags=FLAG_ASYNC)
    b.sync()
    conv_ms = float(b.conv_times_ms(1)[0])
    lib.rsb_debug_tc_cycles(b._h, 0, out, 32)
    cyc = np.array(out[:], dtype=np.float64)
    tiles = max(cyc[14], 1.0)
    res = {"conv_ms": conv_ms, "tiles_cta0": int(cyc[14]), "kernel_cycles_per_tile": round(cyc[13] / tiles, 1),
           "cycles_per_tile": {NAMES[i]: round(cyc[i] / tiles, 1) for i in list(range(13)) + [15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26]}}
    print(json.dumps(res))
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "tc_profile.json").write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
