#!/usr/bin/env python
"""tools/verify_full_size.py -- sample-level verification of BASELINE configs[1] at FULL size
(1024 stereo streams x 60 s, 44.1 -> 48 kHz, 128 taps, 512-frame calls) on one GPU, as SURVEY.md
8(d) asks: counts for every stream, samples of the streams s = 0 mod 64 (plus the last one)
against the CPU oracle fed with the very same input bytes; reports max |diff|, the number of
samples above 1e-6 and the fraction of bit-identical samples, per kernel.  Prints one JSON line.
(The oracle is the checker here, never the thing measured.)"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib as O  # noqa: E402
from resampler_b200 import Attenuation, FirBatch, Kernel, Latency, _lib  # noqa: E402
from resampler_b200.fir import MEM_DEVICE, DeviceBuffer  # noqa: E402

IN_HZ, OUT_HZ, CH, CALL = 44100, 48000, 2, 512


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
    kernels = [Kernel[k.upper()] for k in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["tensor"])]
    lib = _lib.load()
    frames = int(round(seconds * IN_HZ))
    in_stride = frames * CH
    d_in = DeviceBuffer(0, n * in_stride)
    assert lib.rsb_fill_synthetic(0, d_in.ptr, 0, n, frames, CH, IN_HZ, 0x5EED) == 0
    subset = sorted(set(list(range(0, n, 64)) + [n - 1]))
    t0 = time.time()
    refs = {}
    for s in subset:
        x = d_in.download(in_stride, s * in_stride)
        refs[s] = O.OracleFir(CH, IN_HZ, OUT_HZ, 3, 1).process(x, CALL * CH)
    t_oracle = time.time() - t0
    out = {"workload": f"{n} stereo streams x {seconds:g} s, 44.1->48 kHz, 128 taps, {CALL}-frame calls",
           "streams_compared": len(subset), "oracle_seconds": round(t_oracle, 1), "kernels": {}}
    for kern in kernels:
        b = FirBatch(n, CH, IN_HZ, OUT_HZ, Latency.Sample64, Attenuation.Db90, kernel=kern)
        out_stride = ((int(frames / b.ratio()) + 8) * CH + 3) & ~3
        d_out = DeviceBuffer(0, n * out_stride)
        cons, prod, calls = b.process_ptrs([d_in.ptr + 4 * s * in_stride for s in range(n)],
                                           [in_stride] * n, CALL * CH, 0,
                                           [d_out.ptr + 4 * s * out_stride for s in range(n)],
                                           [out_stride] * n, memspace=MEM_DEVICE)
        ref0 = refs[subset[0]]
        counts_ok = all(c == ref0["consumed_total"] for c in cons[:]) and \
            all(p == len(ref0["out"]) for p in prod[:]) and all(k == ref0["calls"] for k in calls[:])
        worst, above, same, total = 0.0, 0, 0, 0
        for s in subset:
            got = d_out.download(prod[s], s * out_stride)
            ref = refs[s]["out"]
            d = np.abs(got.astype(np.float64) - ref)
            worst = max(worst, float(d.max()))
            above += int((d > 1e-6).sum())
            same += int((got.view(np.uint32) == ref.view(np.uint32)).sum())
            total += got.size
        out["kernels"][b.last_kernel().name] = {
            "counts_and_calls_equal_for_all_streams": bool(counts_ok),
            "calls_per_stream": int(calls[0]), "produced_values_per_stream": int(prod[0]),
            "samples_compared": total, "max_abs_diff": worst, "above_1e-6": above,
            "fraction_bit_identical": round(same / total, 6)}
        d_out.free()
        b.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
