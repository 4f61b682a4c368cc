"""Throughput of the batched FFT resampler.  python tools/fft_bench.py [streams] [seconds] [in_hz] [out_hz] [channels]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from resampler_b200 import _lib  # noqa: E402

a = sys.argv[1:]
n = int(a[0]) if a else 1024
secs = float(a[1]) if len(a) > 1 else 10.0
r = bench.fft_leg(_lib.load(), 0, 0, None, n, secs, 3)
print(r["value"], "Msamples/s", r["ms_per_step"], "ms/step", r["roofline"]["frac"])
