"""Error anatomy of the tensor kernel: TENSOR and EXACT outputs against an f64 evaluation of the
reference's formula (same f32 table, same plan), for full-scale uniform noise.  Reports bias
(signed, relative to the output's sign), rms and max.  Under gpurun: python tools/tc2_error_stats.py"""
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from resampler_b200 import Attenuation, FirBatch, Kernel, Latency, _lib  # noqa: E402
from resampler_b200.fir import FLAG_KEEP_PLAN  # noqa: E402


def table(in_hz, out_hz, lat, att, taps):
    lib = _lib.load()
    buf = np.zeros(1024 * taps, np.float32)
    bits = C.c_uint32(0)
    rc = lib.rsb_host_design_table(in_hz, out_hz, lat, att, buf.ctypes.data_as(C.POINTER(C.c_float)),
                                   buf.size, C.byref(bits))
    assert rc == 0
    return buf.reshape(1024, taps).astype(np.float64)


def main():
    ch, in_hz, out_hz, lat, call, n, frames = 2, 44100, 48000, 3, 512, 64, 44100
    if len(sys.argv) > 1:
        in_hz, out_hz = int(sys.argv[1]), int(sys.argv[2])
    taps = 16 << lat
    rng = np.random.default_rng(99)
    xs = [rng.uniform(-1, 1, frames * ch).astype(np.float32) for _ in range(n)]
    T = table(in_hz, out_hz, lat, 1, taps)
    outs, plan = {}, None
    for kern in (Kernel.EXACT, Kernel.TENSOR, Kernel.FAST):
        b = FirBatch(n, ch, in_hz, out_hz, Latency(lat), Attenuation.Db90, kernel=kern)
        res = b.process(xs, call * ch, 0, flags=FLAG_KEEP_PLAN)
        outs[kern.name] = [np.array(o, copy=True) for o in res["out"]]
        if plan is None:
            plan = {k: np.array(v, copy=True) for k, v in b.last_plan(0).items()}
        b.close()
    p1 = plan["phase1"].astype(np.int64)
    p2 = np.minimum(p1 + 1, 1023)
    fr = plan["frac_bits"].view(np.float32)
    omf = (np.float32(1.0) - fr).astype(np.float64)
    fr64 = fr.astype(np.float64)
    k = np.arange(len(p1), dtype=np.float64)
    v = np.rint(k * in_hz / out_hz - (p1 + fr64) / 1024.0).astype(np.int64)
    n_use = 8
    res = {}
    true_all = []
    for s in range(n_use):
        x = xs[s].astype(np.float64).reshape(frames, ch)
        idx = v[:, None] + np.arange(taps)[None, :]
        acc = np.zeros((len(p1), ch))
        for c in range(ch):
            w = x[:, c][idx]
            acc[:, c] = (w * T[p1]).sum(1) * omf + (w * T[p2]).sum(1) * fr64
        true_all.append(acc.reshape(-1))
    true = np.concatenate(true_all)
    for name in outs:
        got = np.concatenate([outs[name][s].astype(np.float64) for s in range(n_use)])
        d = got - true
        big = np.abs(true) > 0.25
        signed = d * np.sign(true)
        res[name] = {"max_abs_err_vs_f64": float(np.abs(d).max()), "rms": float(np.sqrt((d * d).mean())),
                     "mean_signed_err_toward_larger_magnitude": float(signed[big].mean()),
                     "mean_rel_signed_err_in_2^-24": float((signed[big] / np.abs(true[big])).mean() * 2 ** 24),
                     "n": int(d.size)}
    for name in ("TENSOR", "FAST"):
        a = np.concatenate([outs[name][s].astype(np.float64) for s in range(n)])
        e = np.concatenate([outs["EXACT"][s].astype(np.float64) for s in range(n)])
        res[name]["max_abs_diff_vs_EXACT_all_streams"] = float(np.abs(a - e).max())
        res[name]["n_all"] = int(a.size)
    print(json.dumps(res, indent=1))
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / f"tc2_error_stats_{in_hz}_{out_hz}.json").write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
