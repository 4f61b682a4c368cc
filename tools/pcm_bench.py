#!/usr/bin/env python
"""tools/pcm_bench.py -- measurement of the CLI batch path.

Workload: BASELINE configs[1] as the CLI would see it -- 1024 stereo 16-bit PCM files x 60 s,
44.1 -> 48 kHz, 128 taps, 512-value calls (resample/src/main.rs:226-254) -- through
rsb_fir_process_pcm_batch.  Prints ONE JSON line:
  * device-resident: whole-step Msamples/s (format step + plan + convolution + state), the
    format-step kernel's own time and its HBM roofline (algorithmic bytes = raw bytes read +
    f32 bytes written, against the measured copy bandwidth);
  * e2e: the same with pinned HOST buffers (raw s16 up, f32 down), time-sliced over 4 host
    threads like bench.py's e2e leg.
Not a bench.py replacement: BASELINE's metric stays on f32 input.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

IN_HZ, OUT_HZ, CH = 44100, 48000, 2
CALL_LEN = 512   # values, main.rs:227


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=1024)
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--src-channels", type=int, default=2)
    ap.add_argument("--format", default="s16", choices=["s16", "s24"])
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    from resampler_b200 import Attenuation, FirBatch, Latency, PcmFormat, _lib
    from resampler_b200.fir import FLAG_ASYNC, MEM_DEVICE, MEM_HOST
    lib = _lib.load()
    n, sc = args.streams, args.src_channels
    frames = int(round(args.seconds * IN_HZ))
    rng = np.random.default_rng(1)
    t = np.arange(frames, dtype=np.float64)
    one = np.empty((frames, sc), np.int16)
    for c in range(sc):
        sig = 0.5 * np.sin(2 * np.pi * (440 + 7 * c) * t / IN_HZ) + 0.25 * rng.uniform(-1, 1, frames)
        one[:, c] = np.round(sig * 32767).astype(np.int16)
    fmt = PcmFormat.S16
    if args.format == "s24":       # the same signal as packed little-endian 24-bit samples
        fmt = PcmFormat.S24
        v = (one.astype(np.int32) << 8) & 0xFFFFFF
        one = np.stack([v & 0xFF, (v >> 8) & 0xFF, (v >> 16) & 0xFF], axis=-1).astype(np.uint8)
    raw_bytes = one.nbytes
    stride = (raw_bytes + 15) & ~15
    batch = FirBatch(n, CH, IN_HZ, OUT_HZ, Latency.Sample64, Attenuation.Db90)
    cap = (int(frames / batch.ratio()) + 8) * CH
    cap = (cap + 3) & ~3
    d_raw = lib.rsb_alloc_device(0, stride * n)
    d_out = lib.rsb_alloc_device(0, cap * 4 * n)
    assert d_raw and d_out, _lib.last_error()
    assert lib.rsb_memcpy(0, d_raw, one.ctypes.data, raw_bytes, 0) == 0
    for s in range(1, n):
        assert lib.rsb_memcpy(0, d_raw + s * stride, d_raw, raw_bytes, 2) == 0
    in_ptrs = (C.c_void_p * n)(*[d_raw + s * stride for s in range(n)])
    out_ptrs = (C.c_void_p * n)(*[d_out + s * cap * 4 for s in range(n)])
    fr = (C.c_size_t * n)(*[frames] * n)
    caps = (C.c_size_t * n)(*[cap] * n)

    def step():
        batch.reset(-1)
        return batch.process_pcm_ptrs(in_ptrs, fr, fmt, sc, CALL_LEN, 0, out_ptrs, caps,
                                      memspace=MEM_DEVICE, flags=FLAG_ASYNC)

    import os

    def measure():
        for _ in range(args.warmup):
            step()
        ingest, conv = [], []
        batch.timer_start()
        for _ in range(args.steps):
            _, prod, _ = step()
        ms = batch.timer_stop() / args.steps
        batch.sync()
        ingest.append(batch.last_ingest_ms())            # of the last step
        conv = [float(x) for x in batch.conv_times_ms(args.steps)]
        return ms, int(sum(prod[:])), float(np.mean(ingest)), float(np.mean(conv)), batch.last_pcm_fused()

    # (1) as shipped: stereo s16 at an aligned stride => the tensor kernel converts in its loader
    ms, produced, tail_ms, conv_ms, fused = measure()
    # (2) the separate format pass over the whole input (what every other format / layout runs)
    os.environ["RSB_PCM_UNFUSED"] = "1"
    ms_u, produced_u, ingest_ms, conv_ms_u, fused_u = measure()
    del os.environ["RSB_PCM_UNFUSED"]
    assert produced == produced_u and not fused_u
    ingest_bytes = n * (raw_bytes + frames * CH * 4)
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    hbm = float(peaks.get("hbm_gbs", 6554.2))
    # algorithmic HBM bytes of the fused convolution: raw frames in (2 B per value), f32 out
    conv_bytes = n * raw_bytes + produced * 4
    res = {
        "workload": f"{n} files x {args.seconds:g} s, {args.format} PCM {sc} ch -> stereo f32, 44.1->48 kHz, "
                    f"128 taps, {CALL_LEN}-value calls, device-resident, inputs larger than L2",
        "metric": "output Msamples/s (CLI batch path: format step + FIR)",
        "value": round(produced / ms / 1e3, 3), "unit": "Msamples/s", "ms_per_step": round(ms, 4),
        "steps": args.steps, "warmup": args.warmup, "kernel": batch.last_kernel().name,
        "format_step_fused_into_conv": bool(fused), "conv_ms": round(conv_ms, 4),
        "history_tail_ingest_ms": round(tail_ms, 4),
        "roofline_conv": {"bound": "hbm", "kernel": f"conv_tc_kernel<2, raw {args.format}>",
                          "achieved": round(conv_bytes / conv_ms / 1e6, 1), "peak": hbm, "unit": "GB/s",
                          "frac": round(conv_bytes / conv_ms / 1e6 / hbm, 4),
                          "algorithmic": "raw input bytes + 4 B per output value"},
        "separate_format_pass": {
            "value": round(produced_u / ms_u / 1e3, 3), "ms_per_step": round(ms_u, 4),
            "ingest_ms": round(ingest_ms, 4), "conv_ms": round(conv_ms_u, 4),
            "roofline_ingest": {"bound": "hbm", "kernel": f"pcm_ingest_kernel<{args.format}>",
                                "achieved": round(ingest_bytes / ingest_ms / 1e6, 1), "peak": hbm,
                                "unit": "GB/s", "frac": round(ingest_bytes / ingest_ms / 1e6 / hbm, 4),
                                "algorithmic": "raw bytes read + 4 B per converted value written"}},
        "gpu_launches": int(batch.launch_count()),
    }
    lib.rsb_free_device(0, d_raw)
    lib.rsb_free_device(0, d_out)
    batch.close()

    if not args.no_e2e and args.format == "s16":
        res["e2e"] = e2e(lib, n, frames, sc, one)
    print(json.dumps(res))


def e2e(lib, n, frames, sc, one):
    """Pinned host buffers: raw s16 host->device and f32 results device->host inside the timed
    region; 2 s time slices, 4 host threads each driving its own handle."""
    from resampler_b200 import Attenuation, FirBatch, Latency, PcmFormat
    from resampler_b200.fir import MEM_HOST
    n_workers = 4
    slice_frames = 2 * IN_HZ - (2 * IN_HZ) % (CALL_LEN // CH)
    n_slices = max(1, frames // slice_frames)
    in_bytes = slice_frames * sc * 2
    in_stride = (in_bytes + 15) & ~15
    out_vals = (int(slice_frames * OUT_HZ / IN_HZ) + 4400) * CH
    h_in = lib.rsb_alloc_pinned(n * in_stride)
    h_out = lib.rsb_alloc_pinned(n * out_vals * 4)
    assert h_in and h_out
    src = np.ctypeslib.as_array(C.cast(h_in, C.POINTER(C.c_uint8)), shape=(n, in_stride))
    src[:, :in_bytes] = one[:slice_frames].reshape(-1).view(np.uint8)[None, :]
    bounds = [n * w // n_workers for w in range(n_workers + 1)]
    workers = []
    for w in range(n_workers):
        lo, hi = bounds[w], bounds[w + 1]
        fb = FirBatch(hi - lo, CH, IN_HZ, OUT_HZ, Latency.Sample64, Attenuation.Db90)
        cnt = hi - lo
        workers.append((fb, (C.c_void_p * cnt)(*[h_in + s * in_stride for s in range(lo, hi)]),
                        (C.c_void_p * cnt)(*[h_out + 4 * s * out_vals for s in range(lo, hi)]),
                        (C.c_size_t * cnt)(*[slice_frames] * cnt), (C.c_size_t * cnt)(*[out_vals] * cnt)))
    produced = [0] * n_workers

    def work(w, n_steps):
        fb, ip, op, fr, caps = workers[w]
        tot = 0
        for _ in range(n_steps):
            fb.reset(-1)
            for _ in range(n_slices):
                _, p, _ = fb.process_pcm_ptrs(ip, fr, PcmFormat.S16, sc, CALL_LEN, 0, op, caps,
                                              memspace=MEM_HOST)
                tot += int(sum(p[:]))
        produced[w] = tot

    def run(n_steps):
        ths = [threading.Thread(target=work, args=(w, n_steps)) for w in range(n_workers)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        return sum(produced)

    run(1)
    steps = 2
    t0 = time.perf_counter()
    total = run(steps)
    dt = time.perf_counter() - t0
    for fb, *_ in workers:
        fb.close()
    lib.rsb_free_pinned(h_in)
    lib.rsb_free_pinned(h_out)
    return {"value": round(total / dt / 1e6, 3), "unit": "Msamples/s", "steps": steps,
            "h2d_bytes_per_step": int(n_slices * n * in_bytes),
            "d2h_bytes_per_step": int(total // steps * 4),
            "how": f"rsb_fir_process_pcm_batch(memspace=HOST, pinned), {n_slices} slices of "
                   f"{slice_frames / IN_HZ:.2f} s, {n_workers} host threads / handles, wall clock"}


if __name__ == "__main__":
    main()
