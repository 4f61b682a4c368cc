"""Debug aid: runs the TENSOR kernel next to the EXACT kernel on the same input and prints where
and by how much they differ.  usage: python tools/tc_debug.py [channels] [streams] [frames] [in_hz] [out_hz] [latency]"""
import sys
import numpy as np

sys.path.insert(0, ".")
from resampler_b200.fir import Attenuation, FirBatch, Kernel, Latency  # noqa: E402


def main():
    a = [int(x) for x in sys.argv[1:]]
    ch = a[0] if len(a) > 0 else 2
    n = a[1] if len(a) > 1 else 70
    frames = a[2] if len(a) > 2 else 11000
    in_hz = a[3] if len(a) > 3 else 44100
    out_hz = a[4] if len(a) > 4 else 48000
    lat = a[5] if len(a) > 5 else 3
    call = a[6] if len(a) > 6 else 512
    rng = np.random.default_rng(1)
    xs = [(rng.random(frames * ch, dtype=np.float32) * 2 - 1) for _ in range(n)]
    outs = {}
    for kern in (Kernel.EXACT, Kernel.TENSOR):
        b = FirBatch(n, ch, in_hz, out_hz, Latency(lat), Attenuation.Db90, kernel=kern)
        for rep in range(2):
            res = b.process(xs, call * ch, 0)
            outs[(kern, rep)] = [np.array(o, copy=True) for o in res["out"]]
            print(kern.name, "batch", rep, "ran on", b.last_kernel().name, "produced", res["produced"][0],
                  flush=True)
        b.close()
    for rep in range(2):
        worst = 0.0
        for s in range(n):
            e, t = outs[(Kernel.EXACT, rep)][s], outs[(Kernel.TENSOR, rep)][s]
            assert len(e) == len(t), (len(e), len(t))
            d = np.abs(e.astype(np.float64) - t.astype(np.float64))
            if d.size and d.max() > worst:
                worst = float(d.max())
                where = (s, int(d.argmax()) // ch, int(d.argmax()) % ch)
        print(f"batch {rep}: max |tensor - exact| = {worst:.3e} at (stream, frame, ch) {where}")
        e, t = outs[(Kernel.EXACT, rep)][0], outs[(Kernel.TENSOR, rep)][0]
        d = np.abs(e.astype(np.float64) - t.astype(np.float64)).reshape(-1, ch).max(axis=1)
        bad = np.nonzero(d > 1e-6)[0]
        print("  stream 0: frames over 1e-6:", len(bad), "of", len(d), "first", bad[:10], "last", bad[-5:])
        if len(bad):
            f = bad[0]
            print("  exact ", e.reshape(-1, ch)[f:f + 4].ravel())
            print("  tensor", t.reshape(-1, ch)[f:f + 4].ravel())


if __name__ == "__main__" and not (len(sys.argv) > 1 and sys.argv[1] == "structure"):
    main()


def structure():
    """Prints the structure of large mismatches: which streams, which frame ranges."""
    a = [int(x) for x in sys.argv[2:]]
    ch, n, frames, in_hz, out_hz, lat, call = a[:7]
    rng = np.random.default_rng(1)
    xs = [(rng.random(frames * ch, dtype=np.float32) * 2 - 1) for _ in range(n)]
    b = FirBatch(n, ch, in_hz, out_hz, Latency(lat), Attenuation.Db90, kernel=Kernel.EXACT)
    want = [np.array(o, copy=True) for o in b.process(xs, call * ch, 0)["out"]]
    b.close()
    for attempt in range(int(__import__("os").environ.get("TC_ATTEMPTS", "80"))):
        b = FirBatch(n, ch, in_hz, out_hz, Latency(lat), Attenuation.Db90, kernel=Kernel.TENSOR)
        got = b.process(xs, call * ch, 0)["out"]
        b.close()
        bad = {}
        for s in range(n):
            d = np.abs(got[s].astype(np.float64) - want[s]).reshape(-1, ch).max(axis=1)
            idx = np.nonzero(d > 1e-3)[0]
            if len(idx):
                bad[s] = idx
        if bad:
            streams = sorted(bad)
            print("attempt", attempt, "bad streams", len(streams), "first", streams[:8], "last", streams[-4:])
            s0 = streams[0]
            idx = bad[s0]
            print("  stream", s0, "bad frames", len(idx), "range", idx[0], idx[-1], "tiles", sorted(set(idx // 32))[:10])
            print("  frames in tile:", sorted(set(idx % 32)))
            allt = sorted(set(int(t) for s in streams for t in set(bad[s] // 32)))
            print("  all bad tiles", allt[:20], "n", len(allt))
            rows = sorted(set(s % 128 for s in streams))
            print("  rows within group:", rows[:40], "n", len(rows))
            for s in streams[:3]:
                idx = bad[s]
                dd = np.abs(got[s].astype(np.float64) - want[s])
                pk = int(np.argmax(dd))
                ratio = in_hz / out_hz
                print("   stream", s, "bad outputs", idx[0], "..", idx[-1], "peak at", pk, "err", dd[pk],
                      "-> input frame ~", pk * ratio, "mod16", (pk * ratio) % 16)
            return
    print("no mismatch in 80 attempts")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "structure":
    structure()
