"""End-to-end (pinned host buffers) throughput of rsb_fir_process_batch for the headline batch shape.
   python tools/e2e_probe.py [seconds_per_call] [calls] [streams]"""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from resampler_b200 import Attenuation, FirBatch, Latency, _lib  # noqa: E402
from resampler_b200.fir import FLAG_ASYNC, MEM_HOST  # noqa: E402


ASYNC = True


def main():
    global ASYNC
    import os
    ASYNC = not os.environ.get('E2E_SYNC')
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 15.0
    calls = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    ch, in_hz, out_hz = 2, 44100, 48000
    lib = _lib.load()
    frames = int(secs * in_hz) // 512 * 512
    in_vals = frames * ch
    out_vals = (int(frames * out_hz / in_hz) + 4400) * ch
    h_in = lib.rsb_alloc_pinned(n * in_vals * 4)
    h_out = lib.rsb_alloc_pinned(n * out_vals * 4)
    src = np.ctypeslib.as_array(C.cast(h_in, C.POINTER(C.c_float)), shape=(n, in_vals))
    src[:] = (np.random.default_rng(0).random(in_vals, dtype=np.float32) - 0.5)[None, :]
    fb = FirBatch(n, ch, in_hz, out_hz, Latency.Sample64, Attenuation.Db90)
    ip = [h_in + 4 * s * in_vals for s in range(n)]
    op = [h_out + 4 * s * out_vals for s in range(n)]

    def run():
        fb.reset(-1)
        tot = 0
        for _ in range(calls):
            _, p, _ = fb.process_ptrs(ip, [in_vals] * n, 512 * ch, 0, op, [out_vals] * n, memspace=MEM_HOST,
                                      flags=FLAG_ASYNC if ASYNC else 0)
            tot += int(sum(p[:]))
        fb.sync()
        return tot
    run()
    t0 = time.perf_counter()
    tot = run()
    dt = time.perf_counter() - t0
    b, s = fb.host_pipeline_stats()
    pc = (C.c_double * 2)()
    lib.rsb_pcie_probe(0, 512 << 20, 6, 2, pc)
    floor = max(calls * n * in_vals * 4 / (pc[0] * 1e9), tot * 4 / (pc[1] * 1e9))
    print(f"   duplex pinned copy rates up {pc[0]:.1f} / down {pc[1]:.1f} GB/s -> ceiling {tot / floor / 1e6:.0f} Msamples/s, "
          f"achieved {tot / dt / 1e6 / (tot / floor / 1e6):.3f} of it")
    print(f"{secs:g} s x {calls} calls, {n} streams: {tot / dt / 1e6:9.1f} Msamples/s  ({dt * 1e3:.1f} ms; "
          f"up {calls * n * in_vals * 4 / dt / 1e9:.1f} GB/s, down {tot * 4 / dt / 1e9:.1f} GB/s; "
          f"pipeline batches {b} slices {s})")


if __name__ == "__main__":
    main()
