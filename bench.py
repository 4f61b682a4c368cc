#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

metric : output Msamples/s, batched FIR 44.1->48 kHz, 128 taps (Latency::Sample64, Db90)
config : configs[1] = 1024 stereo streams x 60 s synthetic audio per GPU, 512-frame calls
         (5168 virtual resample() calls per stream, all inside one rsb_fir_process_batch)
step   : one pass of the hot path over that whole batch (reset + plan + convolve + state)

  python bench.py --gpus N --steps K --warmup W            our CUDA path
  python bench.py --impl reference ...                     the reference's CPU path (oracle port,
                                                           multi-threaded, one stream per thread)
Under torchrun (N > 1) every rank drives one GPU with its own 1024 streams (weak scaling,
streams are independent: no collective on the data path); rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

IN_HZ, OUT_HZ, CHANNELS, TAPS = 44100, 48000, 2, 128
LATENCY, ATTENUATION = 3, 1          # Sample64, Db90
CALL_FRAMES = 512
METRIC = "output Msamples/s, batched FIR 44.1->48k 128-tap"
FP32_NOMINAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # 74.4: 148 SM x 128 FMA/clk x 1.965 GHz


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=1024, help="streams per GPU")
    ap.add_argument("--seconds", type=float, default=60.0, help="audio seconds per stream")
    ap.add_argument("--kernel", default="auto", choices=["auto", "exact", "fast", "tensor"])
    ap.add_argument("--channels", type=int, default=2, help="experiment only: channels per stream")
    ap.add_argument("--in-hz", type=int, default=44100, help="experiment only")
    ap.add_argument("--out-hz", type=int, default=48000, help="experiment only")
    ap.add_argument("--latency", type=int, default=3, help="experiment only: 0..3 = 16..128 taps")
    ap.add_argument("--call-frames", type=int, default=512, help="experiment only")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-legs", action="store_true", help="skip the short legs of the other configs")
    ap.add_argument("--e2e-slice-seconds", type=float, default=15.0,
                    help="audio seconds per host-memspace call of the e2e leg (pinned buffer size)")
    return ap.parse_args()


def measured_bf16():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 2250.0)))
    return 2250.0


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []          # (host time, csv line)
        self.t_begin = None
        self.t_end = None

    def begin(self):
        """Marks the start of the timed region (the sampler itself is started earlier, before the
        warm-up, because nvidia-smi needs a few hundred ms before its first line)."""
        self.t_begin = time.time()

    def end(self):
        self.t_end = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "50", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lines = self.lines
        if self.t_begin is not None and self.t_end is not None:
            inside = [ln for t, ln in lines if self.t_begin <= t <= self.t_end + 0.05]
            # a very short timed region may fall between two samples: then the nearest ones
            if not inside and lines:
                mid = 0.5 * (self.t_begin + self.t_end)
                inside = [ln for _, ln in sorted(lines, key=lambda x: abs(x[0] - mid))[:2]]
        else:
            inside = [ln for _, ln in lines]
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(smax) if smax else None, "samples": len(sm),
                "reasons": sorted(reasons)}


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
    return rank, world, local, dist


def barrier(dist, local):
    if dist is not None:
        import torch
        dist.barrier()
        torch.cuda.synchronize(local)


def reduce_max(dist, local, value: float) -> float:
    from resampler_b200.sharding import all_reduce_scalar
    return all_reduce_scalar(dist, value, "max", f"cuda:{local}" if dist is not None else None)


def reduce_sum(dist, local, value: float) -> float:
    from resampler_b200.sharding import all_reduce_scalar
    return all_reduce_scalar(dist, value, "sum", f"cuda:{local}" if dist is not None else None)


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's multi-threaded AVX-512 port
# ---------------------------------------------------------------------------------------------
def cpu_reference_run(inp: np.ndarray, frames: int, threads: int, repeats: int = 1):
    """inp: [n_streams, frames*CHANNELS].  Returns (Msamples/s, produced per pass, seconds)."""
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_lib as O
    best = None
    produced_total = 0
    for _ in range(repeats):
        secs, produced, _ = O.cpu_bench(inp.shape[0], CHANNELS, IN_HZ, OUT_HZ, LATENCY, ATTENUATION,
                                        inp, frames, CALL_FRAMES, threads, use_avx512=True)
        produced_total = int(produced.sum())
        best = secs if best is None else min(best, secs)
    return produced_total / best / 1e6, produced_total, best


def host_synthetic(n_streams: int, frames: int, seed: int = 0x5EED) -> np.ndarray:
    """Host-side stand-in for the device generator (same spectrum: sine + uniform noise)."""
    rng = np.random.default_rng(seed)
    n = np.arange(frames, dtype=np.float64)
    out = np.empty((n_streams, frames * CHANNELS), np.float32)
    for s in range(n_streams):
        for c in range(CHANNELS):
            f = 110.0 * (1 + (s % 64)) + 7.0 * c
            out[s, c::CHANNELS] = (0.5 * np.sin(2 * np.pi * f * n / IN_HZ) +
                                   0.25 * rng.uniform(-1, 1, frames)).astype(np.float32)
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_lib as O
    threads = os.cpu_count() or 1
    # bounded sample of the same workload: 4 streams per thread x 10 s of audio per step
    n_streams = 4 * threads
    frames = IN_HZ * 10
    inp = host_synthetic(n_streams, frames)
    for _ in range(args.warmup):
        cpu_reference_run(inp, frames, threads)
    t_total, produced = 0.0, 0
    for _ in range(args.steps):
        _, p, secs = cpu_reference_run(inp, frames, threads)
        t_total += secs
        produced += p
    value = produced / t_total / 1e6
    sample = (f"{n_streams} stereo streams x 10 s per step, 512-frame calls, one stream per thread, "
              f"AVX-512 intrinsics={'yes' if O.lib().orc_cpu_has_avx512f() else 'no (scalar-coded 16-lane order)'}")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": "Msamples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(1e3 * t_total / max(args.steps, 1), 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: 1024 stereo streams x 60 s, 44.1->48 kHz, Sample64 "
                               "(128 taps), Db90, 512-frame calls (bounded CPU sample of it)"},
        "cpu_baseline": {"value": round(value, 3), "unit": "Msamples/s", "cores": threads,
                         "per_core": round(value / threads, 2), "cpu_model": cpu_model(),
                         "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 3), "unit": "Msamples/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "note": "the Rust reference cannot be built here (no rustc); this is the line-faithful C "
                "port of its AVX-512 path (oracle/), multi-threaded on the host cores; every step is a "
                "BOUNDED SAMPLE of the configuration (same rates / taps / call size, fewer streams and "
                "seconds), the metric is a rate",
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
# Extra legs (BASELINE.json configs[0], [2], [3] as batched GPU workloads; shorter than the headline
# so that the default run stays within minutes; same streams / channels / rates / taps / call sizes)
EXTRA_LEGS = {
    "C1": dict(name="configs[0] parameters, batched: 1024 stereo streams x 30 s, 48->44.1 kHz, Sample64 "
                    "(128 taps), Db90, 512-frame calls", ch=2, in_hz=48000, out_hz=44100, lat=3,
               streams=1024, seconds=30.0, call=512),
    "C3i": dict(name="configs[2] (i): 4096 mono streams x 30 s, 16->48 kHz, Sample16 (32 taps), Db90, "
                     "160-frame calls", ch=1, in_hz=16000, out_hz=48000, lat=1, streams=4096, seconds=30.0,
                call=160),
    "C4": dict(name="configs[3]: 512 streams x 8 channels x 15 s, 96->48 kHz, Sample32 (64 taps), Db90, "
                    "512-frame calls", ch=8, in_hz=96000, out_hz=48000, lat=2, streams=512, seconds=15.0,
               call=512),
}


def source_stamp():
    """sha256 over the tensor kernel's sources with // comments and blank lines stripped: ties
    profiles/conv_traffic.json to the CODE it was measured on (editing a comment keeps the stamp)."""
    import hashlib
    h = hashlib.sha256()
    for f in ("fir_tc2.cu", "sm100_ptx.cuh", "fir_kernels.h"):
        for line in (ROOT / "resampler_b200" / "csrc" / f).read_text().splitlines():
            code = line.split("//", 1)[0].rstrip()
            if code.strip():
                h.update(code.encode() + b"\n")
    return h.hexdigest()[:16]


def measured_traffic(kernel_name):
    """DRAM bytes of one headline convolution launch, measured by tools/measure_traffic.py (ncu) and
    stamped with the kernel sources' hash; None when the stamp does not match the sources."""
    tp = ROOT / "profiles" / "conv_traffic.json"
    if not tp.exists():
        return None, "no profiles/conv_traffic.json"
    try:
        d = json.loads(tp.read_text()).get(kernel_name, {})
    except Exception:
        return None, "unreadable"
    if d.get("source_stamp") != source_stamp():
        return None, (f"stale: measured for sources {d.get('source_stamp')}, now {source_stamp()} "
                      "(re-run tools/measure_traffic.py)")
    return d.get("dram_bytes_per_launch"), d.get("source", "")


def device_leg(lib, local, rank, cfg, steps, warmup, kern, dist=None, sampler=None):
    """`steps` passes of the hot path over one device-resident batch of the given configuration.
    Returns the per-rank numbers (the caller reduces over ranks)."""
    from resampler_b200 import Attenuation, FirBatch, Latency
    from resampler_b200.fir import FLAG_ASYNC, MEM_DEVICE, DeviceBuffer
    ch, in_hz, out_hz, lat = cfg["ch"], cfg["in_hz"], cfg["out_hz"], cfg["lat"]
    n_streams, call = cfg["streams"], cfg["call"]
    frames = int(round(cfg["seconds"] * in_hz))
    batch = FirBatch(n_streams, ch, in_hz, out_hz, Latency(lat), Attenuation(ATTENUATION), device=local,
                     kernel=kern)
    in_stride = frames * ch
    out_frames_cap = int(frames / batch.ratio()) + 8
    out_stride = (out_frames_cap * ch + 3) & ~3
    d_in = DeviceBuffer(local, n_streams * in_stride)
    d_out = DeviceBuffer(local, n_streams * out_stride)
    rc = lib.rsb_fill_synthetic(local, d_in.ptr, rank * n_streams, n_streams, frames, ch, in_hz, 0x5EED)
    assert rc == 0
    in_ptrs = [d_in.ptr + 4 * s * in_stride for s in range(n_streams)]
    out_ptrs = [d_out.ptr + 4 * s * out_stride for s in range(n_streams)]
    lens, caps = [in_stride] * n_streams, [out_stride] * n_streams

    def step():
        batch.reset(-1)
        return batch.process_ptrs(in_ptrs, lens, call * ch, 0, out_ptrs, caps, memspace=MEM_DEVICE,
                                  flags=FLAG_ASYNC)

    for _ in range(max(warmup, 0)):
        step()
    batch.sync()
    launches0 = batch.launch_count()
    barrier(dist, local)
    if sampler is not None:
        sampler.begin()
    batch.timer_start()
    counts = None
    for _ in range(steps):
        counts = step()
    ms = batch.timer_stop()
    batch.sync()
    if sampler is not None:
        sampler.end()
    barrier(dist, local)
    res = {
        "ms": ms, "launches": batch.launch_count() - launches0,
        "produced": int(sum(counts[1][:])) if counts else 0,
        "consumed": int(sum(counts[0][:])) if counts else 0,
        "calls": int(counts[2][0]) if counts else 0,
        "conv_ms": [float(x) for x in batch.conv_times_ms(min(steps, 64))],
        "kernel": batch.last_kernel().name.lower(), "taps": 16 << lat, "ratio": batch.ratio(),
        "frames": frames, "in_stride": in_stride, "d_in": d_in, "d_out": d_out, "batch": batch,
    }
    return res


def strong_leg(lib, local, rank, world, dist, steps, warmup, kern, total_streams=65536, seconds=10.0):
    """BASELINE configs[4]: 65 536 stereo streams x 10 s, 44.1 -> 48 kHz, 128 taps, sharded BY STREAM ID
    over the N GPUs of the run (strong scaling: the job is fixed, a rank owns streams
    [G*r/N, G*(r+1)/N)).  231 GB in + 252 GB out do not fit two GPUs, so every rank feeds its
    share as time slices of 86 calls (~1 s resident) with the stream state carried from slice to
    slice; the synthetic second of audio is regenerated nowhere: every slice reads the same device
    buffer (the state, the plan and the traffic are those of the real 10 s)."""
    from resampler_b200 import Attenuation, FirBatch, Latency
    from resampler_b200.fir import FLAG_ASYNC, MEM_DEVICE, DeviceBuffer, _ptr_array, _size_array
    from resampler_b200.sharding import shard_range
    lo, hi = shard_range(total_streams, world, rank)
    n = hi - lo
    ch, in_hz, out_hz, lat, call = 2, 44100, 48000, 3, 512
    frames_total = int(round(seconds * in_hz))
    slice_frames = 86 * call                                   # 0.998 s, whole calls
    n_full, tail = divmod(frames_total, slice_frames)
    batch = FirBatch(n, ch, in_hz, out_hz, Latency(lat), Attenuation(ATTENUATION), device=local, kernel=kern)
    in_stride = slice_frames * ch
    out_stride = ((int(slice_frames / batch.ratio()) + 8) * ch + 3) & ~3
    d_in = DeviceBuffer(local, n * in_stride)
    d_out = DeviceBuffer(local, n * out_stride)
    assert lib.rsb_fill_synthetic(local, d_in.ptr, lo, n, slice_frames, ch, in_hz, 0x5EED) == 0
    in_ptrs = _ptr_array([d_in.ptr + 4 * s * in_stride for s in range(n)])
    out_ptrs = _ptr_array([d_out.ptr + 4 * s * out_stride for s in range(n)])
    caps = _size_array([out_stride] * n)
    lens_full, lens_tail = _size_array([in_stride] * n), _size_array([tail * ch] * n)

    def one_pass():
        batch.reset(-1)
        res = []
        for k in range(n_full + (1 if tail else 0)):
            res.append(batch.process_ptrs(in_ptrs, lens_full if k < n_full else lens_tail, call * ch, 0,
                                          out_ptrs, caps, memspace=MEM_DEVICE, flags=FLAG_ASYNC))
        return res

    for _ in range(max(1, min(warmup, 2))):
        one_pass()
    batch.sync()
    barrier(dist, local)
    batch.timer_start()
    res = None
    for _ in range(steps):
        res = one_pass()
    ms = batch.timer_stop()
    batch.sync()
    barrier(dist, local)
    cons = np.sum([np.array(r[0][:], np.int64) for r in res], axis=0)
    prod = np.sum([np.array(r[1][:], np.int64) for r in res], axis=0)
    calls = np.sum([np.array(r[2][:], np.int64) for r in res], axis=0)
    uniform = bool((cons == cons[0]).all() and (prod == prod[0]).all() and (calls == calls[0]).all())
    all_in = bool(cons[0] == frames_total * ch)
    conv_ms = float(np.sum(batch.conv_times_ms(min(len(res), 64))))
    kernel = batch.last_kernel().name.lower()
    d_in.free()
    d_out.free()
    batch.close()
    ms_max = reduce_max(dist, local, ms)
    prod_all = reduce_sum(dist, local, float(prod.sum()))
    cons_all = reduce_sum(dist, local, float(cons.sum()))
    conv_max = reduce_max(dist, local, conv_ms)
    ok_all = reduce_sum(dist, local, 0.0 if (uniform and all_in) else 1.0) == 0.0
    hbm_peak, peak_src = measured_peaks()
    gbs = 4.0 * (prod_all + cons_all) / world / (conv_max * 1e-3) / 1e9
    return {"workload": f"configs[4]: {total_streams} stereo streams x {seconds:g} s, 44.1->48 kHz, Sample64 (128 taps), "
                        f"Db90, 512-frame calls, sharded by stream id over {world} GPU(s) ({n} streams on rank 0), "
                        f"{n_full + (1 if tail else 0)} time slices of <= {slice_frames / in_hz:.3f} s with state carry",
            "scaling": "strong", "kernel": kernel, "value": round(prod_all * steps / (ms_max * 1e-3) / 1e6, 1),
            "unit": "Msamples/s", "ms_per_pass": round(ms_max / steps, 3), "steps": steps,
            "produced_samples_per_pass": int(prod_all), "calls_per_stream": int(calls[0]),
            "counts_uniform_and_complete": ok_all,
            "roofline": {"bound": "hbm", "achieved": round(gbs, 1), "peak": hbm_peak, "unit": "GB/s per GPU",
                         "frac": round(gbs / hbm_peak, 4), "peak_source": peak_src,
                         "conv_ms_per_pass": round(conv_max, 3)}}


def divergent_leg(lib, local, rank, dist, n_streams=4096, submits=64):
    """BASELINE configs[2] variant (ii): 4096 mono streams, 16 -> 48 kHz, Sample16 (32 taps), one
    resample() per stream per submit with PER-STREAM pseudo-random call sizes (80..480 frames): no two
    streams share a plan.  Device-resident, one fused launch per submit (fir_submit.cu); every
    pointer table is built before the timed region."""
    from resampler_b200 import Attenuation, FirBatch, Latency
    from resampler_b200.fir import FLAG_ASYNC, MEM_DEVICE, DeviceBuffer, _ptr_array, _size_array
    ch, in_hz, out_hz, lat = 1, 16000, 48000, 1
    batch = FirBatch(n_streams, ch, in_hz, out_hz, Latency(lat), Attenuation(ATTENUATION), device=local)
    rng = np.random.default_rng(1234 + rank)
    sizes = rng.integers(80, 481, size=(submits, n_streams))
    span = int(sizes.sum(axis=0).max()) + 16
    bso = batch.buffer_size_output()
    d_in = DeviceBuffer(local, n_streams * span)
    d_out = DeviceBuffer(local, n_streams * bso)
    assert lib.rsb_fill_synthetic(local, d_in.ptr, rank * n_streams, n_streams, span, ch, in_hz, 0x5EED) == 0
    cursor = np.zeros(n_streams, np.int64)
    out_ptrs = _ptr_array([d_out.ptr + 4 * s * bso for s in range(n_streams)])
    out_lens = _size_array([bso] * n_streams)
    tables = []
    for k in range(submits):
        tables.append((_ptr_array([int(d_in.ptr + 4 * (s * span + cursor[s])) for s in range(n_streams)]),
                       _size_array(sizes[k].tolist())))
        cursor += sizes[k]

    def one_pass(timed):
        batch.reset(-1)
        res = []
        if timed:
            batch.timer_start()
        for ip, il in tables:
            res.append(batch.submit_ptrs(ip, il, out_ptrs, out_lens, memspace=MEM_DEVICE, flags=FLAG_ASYNC))
        ms = batch.timer_stop() if timed else 0.0
        batch.sync()
        return res, ms

    one_pass(False)
    barrier(dist, local)
    t0 = time.perf_counter()
    res, ms = one_pass(True)
    wall = time.perf_counter() - t0
    cons = np.sum([np.array(r[0][:], np.int64) for r in res], axis=0)
    prod = np.sum([np.array(r[1][:], np.int64) for r in res], axis=0)
    complete = bool((cons == sizes.sum(axis=0)).all())
    kernel = batch.last_kernel().name.lower()
    d_in.free()
    d_out.free()
    batch.close()
    ms_max = reduce_max(dist, local, ms)
    wall_max = reduce_max(dist, local, wall)
    prod_all = reduce_sum(dist, local, float(prod.sum()))
    return {"workload": f"configs[2] (ii): {n_streams} mono streams per GPU, 16->48 kHz, Sample16 (32 taps), Db90, "
                        f"{submits} submits of one resample() per stream, per-stream random call sizes 80..480 frames "
                        "(every stream its own plan), device resident",
            "kernel": kernel + " (fused single-launch submit)", "value": round(prod_all / (ms_max * 1e-3) / 1e6, 1),
            "unit": "Msamples/s", "us_per_submit_device": round(ms_max * 1e3 / submits, 2),
            "us_per_submit_wall": round(wall_max * 1e6 / submits, 2), "submits": submits,
            "all_input_consumed": complete}


def fft_leg(lib, local, rank, dist, n_streams=1024, seconds=10.0, steps=3):
    """The reference's second resampler (`ResamplerFft`) as a batched GPU workload:
    1024 stereo streams x 10 s, 44.1 -> 48 kHz (chunks of 1176 -> 1280 frames), device resident, one
    launch per step.  HBM roofline: 4 B per input value + 4 B per output value."""
    from resampler_b200 import FftBatch
    from resampler_b200.fir import FLAG_ASYNC, MEM_DEVICE, DeviceBuffer, _ptr_array, _size_array
    ch, in_hz, out_hz = 2, 44100, 48000
    b = FftBatch(n_streams, ch, in_hz, out_hz, device=local)
    csi, cso = b.chunk_size_input(), b.chunk_size_output()
    chunks = int(seconds * in_hz * ch) // csi
    d_in = DeviceBuffer(local, n_streams * chunks * csi)
    d_out = DeviceBuffer(local, n_streams * chunks * cso)
    assert lib.rsb_fill_synthetic(local, d_in.ptr, rank * n_streams, n_streams, chunks * csi // ch, ch, in_hz, 0x5EED) == 0
    ip = _ptr_array([d_in.ptr + 4 * s * chunks * csi for s in range(n_streams)])
    op = _ptr_array([d_out.ptr + 4 * s * chunks * cso for s in range(n_streams)])
    il, ol = _size_array([chunks * csi] * n_streams), _size_array([chunks * cso] * n_streams)
    for _ in range(2):
        b.reset(-1)
        b.process_ptrs(ip, il, op, ol, memspace=MEM_DEVICE, flags=FLAG_ASYNC)
    b.sync()
    barrier(dist, local)
    t0 = time.perf_counter()
    for _ in range(steps):
        b.reset(-1)
        done = b.process_ptrs(ip, il, op, ol, memspace=MEM_DEVICE, flags=FLAG_ASYNC)
    b.sync()
    dt = time.perf_counter() - t0
    produced = int(sum(done[:])) * cso
    consumed = int(sum(done[:])) * csi
    d_in.free()
    d_out.free()
    b.close()
    dt_max = reduce_max(dist, local, dt)
    prod_all = reduce_sum(dist, local, float(produced))
    hbm_peak, peak_src = measured_peaks()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    gbs = 4.0 * (produced + consumed) * steps / dt_max / 1e9
    return {"workload": f"ResamplerFft: {n_streams} stereo streams x {seconds:g} s per GPU, 44.1->48 kHz, chunks of "
                        f"{csi // ch} -> {cso // ch} frames ({chunks} per stream), device resident, one launch per step",
            "kernel": "fft_resample_kernel (overlap-add, mixed-radix Stockham FFTs in shared memory)",
            "value": round(prod_all * steps / dt_max / 1e6, 1), "unit": "Msamples/s",
            "ms_per_step": round(dt_max / steps * 1e3, 3), "steps": steps,
            "timing": "wall clock around the asynchronous launches and the final sync (one launch per step)",
            "roofline": {"bound": "hbm", "achieved": round(gbs, 1), "peak": hbm_peak, "unit": "GB/s per GPU",
                         "frac": round(gbs / hbm_peak, 4), "peak_source": peak_src,
                         "note": "compute / shared-memory bound, not HBM bound (DESIGN 3.6)"}}


def leg_roofline(res, hbm_peak, peak_src):
    conv_avg_ms = float(np.mean(res["conv_ms"])) if res["conv_ms"] else float("nan")
    alg_bytes = 4.0 * (res["produced"] + res["consumed"])
    alg_flops = res["produced"] * 2.0 * res["taps"]
    hbm_gbs = alg_bytes / (conv_avg_ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": round(hbm_gbs, 1), "peak": hbm_peak, "unit": "GB/s",
            "frac": round(hbm_gbs / hbm_peak, 4), "peak_source": peak_src,
            "conv_ms_per_launch": round(conv_avg_ms, 4),
            "algorithmic_bytes_per_launch": int(alg_bytes),
            "algorithmic_tflops": round(alg_flops / (conv_avg_ms * 1e-3) / 1e12, 2)}, conv_avg_ms, alg_bytes, alg_flops


def cpu_model():
    try:
        for line in Path("/proc/cpuinfo").read_text().splitlines():
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_ours(args):
    from resampler_b200 import Kernel
    from resampler_b200 import _lib
    import ctypes as C

    all_cpus = os.sched_getaffinity(0)
    numa = numa_affinity(int(os.environ.get("LOCAL_RANK", "0")))     # before the CUDA context exists
    rank, world, local, dist = dist_setup(args.gpus)
    lib = _lib.load()
    kern = {"auto": Kernel.AUTO, "exact": Kernel.EXACT, "fast": Kernel.FAST,
            "tensor": Kernel.TENSOR}[args.kernel]
    head = dict(name="configs[1]", ch=CHANNELS, in_hz=IN_HZ, out_hz=OUT_HZ, lat=LATENCY,
                streams=args.streams, seconds=args.seconds, call=CALL_FRAMES)

    # ---- device-resident throughput of the headline configuration (value) ----
    sampler = ClockSampler(local)
    sampler.start()
    R = device_leg(lib, local, rank, head, args.steps, args.warmup, kern, dist, sampler)
    clocks = sampler.stop()
    batch, d_in = R["batch"], R["d_in"]
    n_streams, frames, in_stride = args.streams, R["frames"], R["in_stride"]
    ms_max = reduce_max(dist, local, R["ms"])
    produced_all = reduce_sum(dist, local, float(R["produced"]))
    value = produced_all * args.steps / (ms_max * 1e-3) / 1e6          # Msamples/s, whole job

    # ---- roofline of the dominant kernel (the convolution) ----
    hbm_peak, peak_src = measured_peaks()
    roof, conv_avg_ms, alg_bytes, alg_flops = leg_roofline(R, hbm_peak, peak_src)
    fp32_tflops = alg_flops / (conv_avg_ms * 1e-3) / 1e12
    fp32_meas = None
    r, a = C.c_double(0), C.c_double(0)
    if rank == 0 and lib.rsb_microbench(local, 1, 4, C.byref(r), C.byref(a)) == 0:
        fp32_meas = r.value
    ran = R["kernel"]          # the kernel that actually served the batches
    traffic, traffic_src = measured_traffic(ran)
    fp32_peak = fp32_meas if fp32_meas else FP32_NOMINAL_TFLOPS
    algorithmic = "2*taps flop and 4*(1+in/out) B per output sample (SURVEY.md 8(d))"
    if ran == "tensor":
        # tcgen05 kernel: the algorithmic HBM floor at the measured copy bandwidth is the roofline
        # the kernel is held against; the tensor pipe is reported beside it.  Executed tensor flops
        # = 3 fp16 products per tap over the tile's padded K range (upper bound kt_max / taps).
        bf16_peak = measured_bf16()
        span = int(63 * R["ratio"]) + 1
        kt_max = (15 + span + R["taps"] + 15) & ~15
        executed = fp32_tflops * 3.0 * kt_max / R["taps"]
        roofline = dict(roof)
        roofline.update({
            "kernel": "conv_tc2_kernel (tcgen05 kind::f16, dominant kernel of the step)",
            "traffic": traffic, "traffic_source": traffic_src,
            "tensor": {"achieved_algorithmic": round(fp32_tflops, 2),
                       "executed_fp16_upper_bound": round(executed, 1),
                       "peak_bf16_measured_sustained": bf16_peak, "unit": "TFLOP/s",
                       "frac_executed_of_fp16_peak": round(executed / bf16_peak, 4) if bf16_peak else None,
                       "note": "3 fp16 products per tap x K padding kt_max/taps; fp16 peak = measured bf16 peak"},
            "fp32_cuda_core_equivalent": {"achieved": round(fp32_tflops, 2), "peak": round(fp32_peak, 2),
                                          "unit": "TFLOP/s",
                                          "note": "the same algorithmic flops against the FFMA2 peak"},
            "algorithmic": algorithmic,
        })
    else:
        roofline = {
            "bound": "fp32", "kernel": "conv (dominant kernel of the step)",
            "achieved": round(fp32_tflops, 3), "peak": round(fp32_peak, 2), "unit": "TFLOP/s",
            "frac": round(fp32_tflops / fp32_peak, 4),
            "peak_source": ("FFMA2 register micro-benchmark run in this process"
                            if fp32_meas else "nominal 148 SM x 128 FMA/clk x 1.965 GHz"),
            "frac_of_nominal_74.4": round(fp32_tflops / FP32_NOMINAL_TFLOPS, 4),
            "traffic": traffic, "traffic_source": traffic_src, "hbm": roof,
            "conv_ms_per_launch": round(conv_avg_ms, 4), "algorithmic": algorithmic,
        }

    # ---- end-to-end through the C ABI with pinned HOST buffers ----
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, batch, lib, frames, n_streams, local, dist)
        e2e["host_affinity"] = numa
    os.sched_setaffinity(0, all_cpus)        # the CPU baseline below uses every core

    # ---- CPU baseline beside it (rank 0, N == 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        cs = min(n_streams, 4 * threads)
        cframes = min(frames, IN_HZ * 10)
        host_in = np.empty((cs, cframes * CHANNELS), np.float32)
        for s in range(cs):      # the very bytes the GPU processed
            host_in[s] = d_in.download(cframes * CHANNELS, s * in_stride)
        est, _, secs1 = cpu_reference_run(host_in, cframes, threads)
        repeats = int(min(max(1, round(12.0 / max(secs1, 1e-3))), 400))
        t_tot, p_tot = 0.0, 0
        for _ in range(repeats):
            _, p, secs = cpu_reference_run(host_in, cframes, threads)
            t_tot += secs
            p_tot += p
        cpu = {"value": round(p_tot / t_tot / 1e6, 3), "unit": "Msamples/s", "cores": threads,
               "per_core": round(p_tot / t_tot / 1e6 / threads, 2), "cpu_model": cpu_model(),
               "kind": "port",
               "sample": f"{cs} of the GPU's own streams x {cframes / IN_HZ:.1f} s, 512-frame calls, "
                         f"one stream per thread, repeated {repeats}x (~{t_tot:.1f} s of CPU work)"}
    # ---- single-stream drop-in call: latency of ONE resample() (512 frames) next to the CPU's ----
    latency = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            from resampler_b200 import Attenuation, Latency, ResamplerFir
            r1 = ResamplerFir.new_from_hz(CHANNELS, IN_HZ, OUT_HZ, Latency(LATENCY), Attenuation(ATTENUATION))
            x = host_synthetic(1, CALL_FRAMES)[0]
            o = np.zeros(r1.buffer_size_output(), np.float32)
            for _ in range(20):
                r1.resample(x, o)
            ts = []
            for _ in range(300):
                t0 = time.perf_counter()
                r1.resample(x, o)
                ts.append(time.perf_counter() - t0)
            r1.close()
            one = host_synthetic(1, IN_HZ * 10)
            _, p1, s1 = cpu_reference_run(one, IN_HZ * 10, 1)
            n_calls_cpu = (IN_HZ * 10 + CALL_FRAMES - 1) // CALL_FRAMES
            latency = {"call": f"rsb_fir_resample, {CALL_FRAMES} stereo frames, host slices, one stream",
                       "gpu_us_median": round(float(np.median(ts)) * 1e6, 1),
                       "gpu_us_p95": round(float(np.percentile(ts, 95)) * 1e6, 1),
                       "cpu_us_per_call": round(s1 / n_calls_cpu * 1e6, 2),
                       "note": "a single stream is latency-bound on a GPU (copy in, 5 launches, copy out); "
                               "the batched entry points are the product, this is the drop-in's cost"}
        except Exception as e:
            latency = {"error": str(e)[:200]}
    head_calls, head_launches = R["calls"], R["launches"]
    d_in.free()
    R["d_out"].free()
    batch.close()

    # ---- the other BASELINE configurations, short legs (device resident, same timing rules) ----
    legs = {}
    if not args.no_legs:
        for key, cfg in EXTRA_LEGS.items():
            try:
                L = device_leg(lib, local, rank, cfg, max(2, min(args.steps, 4)), max(args.warmup, 3), kern, dist)
                lroof, _, _, _ = leg_roofline(L, hbm_peak, peak_src)
                lms = reduce_max(dist, local, L["ms"])
                lprod = reduce_sum(dist, local, float(L["produced"]))
                steps_l = max(2, min(args.steps, 4))
                legs[key] = {"workload": cfg["name"] + " (per GPU)", "kernel": L["kernel"],
                             "value": round(lprod * steps_l / (lms * 1e-3) / 1e6, 1), "unit": "Msamples/s",
                             "ms_per_step": round(lms / steps_l, 4), "steps": steps_l,
                             "calls_per_stream": L["calls"], "roofline": lroof}
                L["d_in"].free()
                L["d_out"].free()
                L["batch"].close()
            except Exception as e:      # a leg must not take the headline line down
                legs[key] = {"workload": cfg["name"], "error": str(e)[:300]}

        try:
            legs["C3ii"] = divergent_leg(lib, local, rank, dist)
        except Exception as e:
            legs["C3ii"] = {"workload": "configs[2] (ii)", "error": str(e)[:300]}
    if not args.no_legs:
        try:
            legs["FFT"] = fft_leg(lib, local, rank, dist)
        except Exception as e:
            legs["FFT"] = {"workload": "ResamplerFft", "error": str(e)[:300]}
    strong = None
    if not args.no_legs:
        try:
            strong = strong_leg(lib, local, rank, world, dist, max(2, min(args.steps, 3)), args.warmup, kern)
        except Exception as e:
            strong = {"workload": "configs[4]", "scaling": "strong", "error": str(e)[:300]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 3), "unit": "Msamples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_max / max(args.steps, 1), 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[1]: {n_streams} stereo streams x {args.seconds:g} s "
                                   f"per GPU, 44.1->48 kHz, Sample64 (128 taps), Db90, 512-frame "
                                   f"virtual calls ({head_calls} per stream)",
                       "kernel": f"{args.kernel} -> {ran}", "streams_per_gpu": n_streams,
                       "l2_policy": "inputs larger than L2 (in+out per step "
                                    f"{(alg_bytes) / 1e9:.1f} GB >> 126 MB)"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(head_launches),
            "clocks": clocks,
            "produced_samples_per_step_per_gpu": R["produced"],
            "other_configs": legs, "strong_scaling": strong, "single_stream_latency": latency,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def numa_affinity(local):
    """Pins this process to the host cores of the GPU's NUMA node (before any pinned allocation, so
    that the staging memory is node-local).  Returns what was done, for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.lower().split(":", 1)
        node = int(Path(f"/sys/bus/pci/devices/{dom[-4:]}:{rest}/numa_node").read_text())
        if node < 0:
            return {"numa_node": node, "pinned": False}
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "pinned": bool(cpus), "cpus": len(cpus)}
    except Exception as e:
        return {"numa_node": None, "pinned": False, "why": str(e)[:80]}


def pcie_ceiling(lib, local, dist):
    """Plain pinned cudaMemcpyAsync rates of this GPU's link (all ranks probe at the same time, so
    at N > 1 these are the contended rates)."""
    import ctypes as C
    out = (C.c_double * 2)()
    res = {}
    for mode, name in ((0, "h2d_alone"), (1, "d2h_alone"), (2, "duplex"), (3, "duplex_2d")):
        barrier(dist, local)
        if lib.rsb_pcie_probe(local, (512 << 20) if mode != 3 else (118 << 20), 6 if mode != 3 else 24, mode, out) != 0:
            return None
        if mode == 2:
            res["h2d_duplex"], res["d2h_duplex"] = round(out[0], 2), round(out[1], 2)
        elif mode == 3:       # 1024 rows of 115 KB: the slices the library's pipeline copies
            res["h2d_duplex_2d_rows"], res["d2h_duplex_2d_rows"] = round(out[0], 2), round(out[1], 2)
        else:
            res[name] = round(out[mode], 2)
    return res


def run_e2e(args, batch, lib, frames, n_streams, local, dist):
    """Same workload through rsb_fir_process_batch with HOST (pinned) buffers: every step copies
    the step's inputs host->device and the results device->host inside the timed region.  One
    handle, one host thread: the library cuts each call into time slices and overlaps the H2D copy
    of slice k+1, the kernels of slice k and the D2H copy of slice k-1 itself.  The 60 s are fed
    as a few calls that reuse one pinned buffer (the state carries over from call to call)."""
    import ctypes as C
    from resampler_b200 import Attenuation, FirBatch, Latency
    from resampler_b200.fir import FLAG_ASYNC, MEM_HOST
    call_seconds = args.e2e_slice_seconds
    part_frames = int(round(call_seconds * IN_HZ))
    part_frames -= part_frames % CALL_FRAMES          # whole calls per part
    part_frames = max(CALL_FRAMES, min(part_frames, frames - frames % CALL_FRAMES))
    n_parts = max(1, frames // part_frames)
    in_vals = part_frames * CHANNELS
    ratio = batch.ratio()
    out_vals = (int(part_frames / ratio) + 4400) * CHANNELS
    pcie = pcie_ceiling(lib, local, dist)
    h_in = lib.rsb_alloc_pinned(n_streams * in_vals * 4)
    h_out = lib.rsb_alloc_pinned(n_streams * out_vals * 4)
    if not h_in or not h_out:
        return {"value": None, "unit": "Msamples/s", "error": "pinned allocation failed"}
    src = np.ctypeslib.as_array(C.cast(h_in, C.POINTER(C.c_float)), shape=(n_streams, in_vals))
    src[:] = host_synthetic(1, part_frames)[0][None, :]
    fb = FirBatch(n_streams, CHANNELS, IN_HZ, OUT_HZ, Latency(LATENCY), Attenuation(ATTENUATION), device=local)
    in_ptrs = [h_in + 4 * s * in_vals for s in range(n_streams)]
    out_ptrs = [h_out + 4 * s * out_vals for s in range(n_streams)]

    def run(n_steps):
        tot = 0
        for _ in range(n_steps):
            fb.reset(-1)
            for _ in range(n_parts):
                _, p, _ = fb.process_ptrs(in_ptrs, [in_vals] * n_streams, CALL_FRAMES * CHANNELS, 0,
                                          out_ptrs, [out_vals] * n_streams, memspace=MEM_HOST, flags=FLAG_ASYNC)
                tot += int(sum(p[:]))
            fb.sync()                        # the step's results are in host memory
        return tot

    steps = max(1, min(args.steps, 3))
    run(1)                                   # warm-up: allocations, first touches
    barrier(dist, local)
    t0 = time.perf_counter()
    total = run(steps)
    dt = time.perf_counter() - t0
    dt_max = reduce_max(dist, local, dt)
    produced_all = reduce_sum(dist, local, float(total))
    pipe_batches, pipe_slices = fb.host_pipeline_stats()
    fb.close()
    lib.rsb_free_pinned(h_in)
    lib.rsb_free_pinned(h_out)
    per_step_p = total // steps
    h2d = int(n_parts * n_streams * in_vals * 4)
    d2h = int(per_step_p * 4)
    res = {"value": round(produced_all / dt_max / 1e6, 3), "unit": "Msamples/s",
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": steps,
           "how": f"rsb_fir_process_batch(memspace=HOST, pinned buffers), {n_parts} calls of "
                  f"{part_frames / IN_HZ:.2f} s per step on one handle / one host thread; inside each call "
                  f"the library pipelines {pipe_slices // max(pipe_batches, 1)} time slices "
                  f"(H2D | kernels | D2H on three streams), calls chained with RSB_FLAG_ASYNC, rsb_fir_sync() per step; "
                  f"wall clock around the calls, max over ranks"}
    if pcie:
        # per-GPU time one step's copies need at the duplex rates: the PCIe floor of a step
        floor_s = max(h2d / (pcie["h2d_duplex"] * 1e9), d2h / (pcie["d2h_duplex"] * 1e9))
        ceiling = per_step_p / floor_s / 1e6
        world = int(os.environ.get("WORLD_SIZE", "1"))
        res["pcie"] = dict(pcie, unit="GB/s per GPU (pinned cudaMemcpyAsync, all ranks at once)",
                           ceiling_msamples_per_gpu=round(ceiling, 1),
                           frac_of_ceiling=round(produced_all / dt_max / 1e6 / (ceiling * world), 4))
    return res


def main():
    global CHANNELS, IN_HZ, OUT_HZ, LATENCY, TAPS, CALL_FRAMES
    args = parse_args()
    CHANNELS = args.channels
    IN_HZ, OUT_HZ, LATENCY, CALL_FRAMES = args.in_hz, args.out_hz, args.latency, args.call_frames
    TAPS = 16 << LATENCY
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
