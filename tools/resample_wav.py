#!/usr/bin/env python
"""tools/resample_wav.py -- the reference CLI's FIR path for MANY files at once.

  python tools/resample_wav.py --sample-rate 48000 [--latency 64] [--attenuation 90] \
         --out-dir OUT  a.wav b.wav ...

Per file this does what `resample --filter fir --sample-rate R in.wav out.wav` does
(resample/src/main.rs:57-216): integer PCM -> f32 by 2^(bits-1), mono duplicated to stereo,
ResamplerFir::new(2, ...) fed in 512-value calls, stereo f32 WAV out -- but all files of one
input rate form ONE GPU batch (one stream per file): raw samples go to the device, the format
step and the resampler run there (rsb_fir_process_pcm_batch).  Defaults as the CLI's:
Latency::Sample64, Attenuation::Db90.  There is no CPU fallback.
"""
from __future__ import annotations

import argparse
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

# main.rs:107-126 (SampleRate::try_from)
SUPPORTED = (16000, 22050, 32000, 44100, 48000, 88200, 96000, 176400, 192000, 384000)


def resample_files(paths, out_dir, sample_rate, latency=64, attenuation=90, device=0, devices=None):
    from resampler_b200 import Attenuation, FirBatch, Latency
    from resampler_b200.wav import read_wav, write_wav_f32
    lat = {8: Latency.Sample8, 16: Latency.Sample16, 32: Latency.Sample32, 64: Latency.Sample64}
    att = {60: Attenuation.Db60, 90: Attenuation.Db90, 120: Attenuation.Db120}
    if latency not in lat:
        raise SystemExit(f"Error: Invalid latency value: {latency}. Must be 8, 16, 32, or 64")
    if attenuation not in att:
        raise SystemExit(f"Error: Invalid attenuation value: {attenuation}. Must be 60, 90, or 120")
    if sample_rate not in SUPPORTED:
        raise SystemExit(f"Unsupported output sample rate: {sample_rate}. Supported rates: "
                         + ", ".join(map(str, SUPPORTED)))
    out_dir = Path(out_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    groups = defaultdict(list)          # (input rate, format, channels) -> files
    names = {}
    for p in paths:
        # every output lands in out_dir under the input's base name: two inputs with the same
        # base name would overwrite each other (possibly from two threads)
        if Path(p).name in names:
            raise SystemExit(f"Error: {p} and {names[Path(p).name]} would both be written to "
                             f"{Path(out_dir) / Path(p).name}")
        names[Path(p).name] = p
        w = read_wav(p)
        if w.sample_rate not in SUPPORTED:
            raise SystemExit(f"Unsupported input sample rate: {w.sample_rate}. Supported rates: "
                             + ", ".join(map(str, SUPPORTED)))
        if w.channels not in (1, 2):
            raise SystemExit(f"Unsupported channel count: {w.channels}")     # main.rs:151-154
        groups[(w.sample_rate, w.fmt, w.channels)].append((Path(p), w))
    # Files are independent streams: every (rate, format, channels) group is cut into one
    # contiguous share per GPU (resampler_b200.sharding), one host thread and handle per share,
    # no collective.
    import threading
    from resampler_b200.sharding import shard_range
    devices = list(devices) if devices else [device]
    written, errors, lock = [], [], threading.Lock()

    def run_share(dev, in_rate, fmt, ch, items):
        try:
            batch = FirBatch(len(items), 2, in_rate, sample_rate, lat[latency], att[attenuation],
                             device=dev)
            res = batch.process_pcm([w.raw for _, w in items], fmt, ch, call_len=512)
            batch.close()
            for (p, w), out in zip(items, res["out"]):
                dst = out_dir / p.name
                write_wav_f32(dst, out, sample_rate, 2)
                with lock:
                    written.append((dst, w.frames, len(out) // 2))
        except Exception as e:            # surfaced after the join
            with lock:
                errors.append(e)

    threads = []
    for (in_rate, fmt, ch), items in groups.items():
        for r, dev in enumerate(devices):
            lo, hi = shard_range(len(items), len(devices), r)
            if hi > lo:
                threads.append(threading.Thread(target=run_share,
                                                args=(dev, in_rate, fmt, ch, items[lo:hi])))
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return sorted(written)


def main():
    ap = argparse.ArgumentParser(description="Resample WAV files to different sample rates (FIR, GPU batch)")
    ap.add_argument("--sample-rate", type=int, required=True, metavar="RATE")
    ap.add_argument("--latency", type=int, default=64, metavar="SAMPLES")
    ap.add_argument("--attenuation", type=int, default=90, metavar="DB")
    ap.add_argument("--out-dir", required=True)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--gpus", type=int, default=1, help="shard the files over this many GPUs")
    ap.add_argument("inputs", nargs="+")
    a = ap.parse_args()
    devices = list(range(a.device, a.device + a.gpus))
    for dst, fin, fout in resample_files(a.inputs, a.out_dir, a.sample_rate, a.latency,
                                         a.attenuation, a.device, devices):
        print(f"{dst}: Input frames: {fin}  Output frames: {fout}")


if __name__ == "__main__":
    main()
