"""Bring-up check of the tensor kernel (fir_tc2.cu): TENSOR against EXACT on the GPU for a list of
configurations; prints max |diff|, where the worst sample sits and the error histogram by
(stream, frame mod 64).  Under gpurun:  python tools/tc2_check.py [quick]"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from resampler_b200 import Attenuation, FirBatch, Kernel, Latency  # noqa: E402

CASES = [
    # ch, in_hz, out_hz, latency, call frames, streams, frames
    (2, 44100, 48000, 3, 512, 64, 3000),
    (2, 44100, 48000, 3, 512, 70, 9000),
    (1, 16000, 48000, 1, 160, 130, 4000),
    (2, 48000, 44100, 3, 512, 9, 9000),
    (8, 96000, 48000, 2, 512, 17, 9000),
    (4, 44100, 48000, 3, 512, 40, 9000),
    (1, 44100, 48000, 3, 512, 200, 9000),
    (2, 44100, 48000, 3, 512, 512, 44100),
]


def run(case, kind):
    ch, in_hz, out_hz, lat, call, n, frames = case
    rng = np.random.default_rng(1234)
    if kind == "noise":
        xs = [rng.uniform(-1, 1, frames * ch).astype(np.float32) for _ in range(n)]
    else:   # impulses: one nonzero sample per stream at a stream-dependent position
        xs = []
        for s in range(n):
            x = np.zeros(frames * ch, np.float32)
            x[((200 + 3 * s) % frames) * ch + (s % ch)] = 1.0
            xs.append(x)
    outs = {}
    for kern in (Kernel.EXACT, Kernel.TENSOR):
        b = FirBatch(n, ch, in_hz, out_hz, Latency(lat), Attenuation.Db90, kernel=kern)
        res = b.process(xs, call * ch, 0)
        outs[kern] = [np.array(o, copy=True) for o in res["out"]]
        ran = b.last_kernel()
        b.close()
        if kern == Kernel.TENSOR and ran != Kernel.TENSOR:
            return {"case": case, "kind": kind, "error": f"ran on {ran.name}"}
    worst, where, nbad, total = 0.0, None, 0, 0
    for s in range(n):
        a, e = outs[Kernel.TENSOR][s].astype(np.float64), outs[Kernel.EXACT][s].astype(np.float64)
        if len(a) != len(e):
            return {"case": case, "kind": kind, "error": f"length {len(a)} vs {len(e)} stream {s}"}
        d = np.abs(a - e)
        d = np.where(np.isfinite(d), d, 1e30)
        total += d.size
        nbad += int((d > 1e-6).sum())
        if d.size and d.max() > worst:
            worst = float(d.max())
            i = int(d.argmax())
            where = {"stream": s, "value_index": i, "frame": i // ch, "frame_mod_64": (i // ch) % 64,
                     "got": float(a[i]), "want": float(e[i])}
    return {"case": case, "kind": kind, "max_abs_diff": worst, "n_over_1e-6": nbad, "n": total,
            "where": where}


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    cases = CASES[:3] if quick else CASES
    results = []
    for case in cases:
        for kind in ("impulse", "noise"):
            r = run(case, kind)
            print(json.dumps(r), flush=True)
            results.append(r)
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "tc2_check.json").write_text(json.dumps(results, indent=1))


if __name__ == "__main__":
    main()
