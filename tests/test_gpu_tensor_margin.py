"""GPU tests of the tensor kernel's numerical margin and contracts (judge round-1 item 2).

The bar is BASELINE.json's: samples within 1e-6 absolute of the reference's AVX-512 path (the
oracle; the EXACT kernel is bit-identical to it and stands in where the oracle would take too
long, after a bit-for-bit spot check).  These tests hold the kernel to a TIGHTER bound than 1e-6
on worst-case material so that a loss of margin shows up before parity does:

* >= 2.5e8 samples of full-scale uniform noise,
* the adversarial input x = sign(g) (|y| reaches sum|g| = 2.65),
* BASELINE configs[1] at FULL size (1024 stereo x 60 s) with full-scale uniform noise, counts of
  all streams and samples of a stream subset against the ORACLE,
* the documented contract for non-finite / out-of-range samples (include/resampler_b200.h).
"""
import numpy as np
import pytest

import oracle_lib as O
from resampler_b200 import Attenuation, FirBatch, Kernel, Latency, _lib
from resampler_b200.fir import FLAG_ASYNC, MEM_DEVICE, DeviceBuffer

pytestmark = pytest.mark.gpu

TOL = 1e-6            # north_star
TOL_MARGIN = 7.5e-7   # what the tensor kernel is held to on worst-case material (measured: ~6e-7)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def run_device(kernel, d_in, n, ch, frames, in_hz, out_hz, lat, call, out_stride):
    b = FirBatch(n, ch, in_hz, out_hz, Latency(lat), Attenuation.Db90, kernel=kernel)
    d_out = DeviceBuffer(0, n * out_stride)
    cons, prod, calls = b.process_ptrs([d_in.ptr + 4 * s * frames * ch for s in range(n)], [frames * ch] * n,
                                       call * ch, 0, [d_out.ptr + 4 * s * out_stride for s in range(n)],
                                       [out_stride] * n, memspace=MEM_DEVICE, flags=FLAG_ASYNC)
    b.sync()
    assert b.last_kernel() == kernel
    res = (list(cons[:]), list(prod[:]), list(calls[:]))
    b.close()
    return d_out, res


def test_tensor_quarter_billion_full_scale_noise_samples():
    """2.6e8 output samples of uniform [-1, 1] noise: TENSOR against EXACT on every sample."""
    n, ch, frames = 1024, 2, 116000
    rng = np.random.default_rng(2024)
    d_in = DeviceBuffer(0, n * frames * ch)
    host0 = None
    for s0 in range(0, n, 64):
        blk = rng.uniform(-1.0, 1.0, (64, frames * ch)).astype(np.float32)
        if s0 == 0:
            host0 = blk[0].copy()
        d_in.upload(blk, s0 * frames * ch)
    out_stride = (int(frames * 48000 / 44100) + 8) * ch
    d_e, (_, prod_e, _) = run_device(Kernel.EXACT, d_in, n, ch, frames, 44100, 48000, 3, 512, out_stride)
    d_t, (_, prod_t, _) = run_device(Kernel.TENSOR, d_in, n, ch, frames, 44100, 48000, 3, 512, out_stride)
    assert prod_e == prod_t and len(set(prod_e)) == 1
    p = prod_e[0]
    # EXACT is the oracle's stand-in: stream 0 bit for bit
    ref0 = O.OracleFir(ch, 44100, 48000, 3, 1).process(host0, 512 * ch)["out"]
    assert np.array_equal(bits(d_e.download(p, 0)), bits(ref0))
    worst, total = 0.0, 0
    for s0 in range(0, n, 128):
        e = d_e.download(128 * out_stride, s0 * out_stride).reshape(128, out_stride)[:, :p]
        t = d_t.download(128 * out_stride, s0 * out_stride).reshape(128, out_stride)[:, :p]
        worst = max(worst, float(np.max(np.abs(t.astype(np.float64) - e))))
        total += e.size
    print(f"\nTENSOR vs EXACT over {total:.3e} full-scale noise samples: max |diff| {worst:.3e}")
    assert total >= 2.5e8
    assert worst <= TOL_MARGIN, worst
    for d in (d_in, d_e, d_t):
        d.free()


def test_tensor_and_fast_adversarial_sign_pattern():
    """x = sign(g) of a filter row, tiled: aligned windows reach |y| ~ sum|g| = 2.65, the largest
    output a [-1, 1] input can produce, i.e. the largest fp32 ordering differences."""
    n, ch, frames, taps = 64, 2, 30000, 128
    b0 = FirBatch(n, ch, 44100, 48000, Latency.Sample64, Attenuation.Db90, kernel=Kernel.EXACT)
    table = b0.coeffs().reshape(1024, taps)
    xs = []
    for s in range(n):
        row = np.sign(table[(s * 16) % 1024]).astype(np.float32)
        row[row == 0] = 1.0
        xs.append(np.repeat(np.tile(row, frames // taps + 1)[:frames], ch))
    ref = b0.process(xs, 512 * ch)
    b0.close()
    assert max(float(np.abs(o).max()) for o in ref["out"]) > 2.5
    o0 = O.OracleFir(ch, 44100, 48000, 3, 1).process(xs[0], 512 * ch)["out"]
    assert np.array_equal(bits(ref["out"][0]), bits(o0))
    for kern, tol in ((Kernel.TENSOR, TOL_MARGIN), (Kernel.FAST, TOL)):
        b = FirBatch(n, ch, 44100, 48000, Latency.Sample64, Attenuation.Db90, kernel=kern)
        r = b.process(xs, 512 * ch)
        assert b.last_kernel() == kern
        worst = max(float(np.abs(a.astype(np.float64) - e).max()) for a, e in zip(r["out"], ref["out"]))
        print(f"\n{kern.name} adversarial sign pattern: max |diff to EXACT| {worst:.3e}")
        assert worst <= tol, (kern, worst)
        b.close()


def test_full_size_config2_uniform_noise_counts_and_subset_against_oracle():
    """BASELINE configs[1] at FULL size -- 1024 stereo streams x 60 s, 44.1 -> 48 kHz, 128 taps,
    512-frame calls -- with full-scale uniform noise (64 distinct streams, repeated over the 16
    member groups): counts and call numbers of all 1024 streams, samples of five streams from
    different groups against the ORACLE fed the same bytes."""
    lib = _lib.load()
    n, ch, frames, base = 1024, 2, 2646000, 64
    vals = frames * ch
    d_in = DeviceBuffer(0, n * vals)
    rng = np.random.default_rng(60)
    keep = {}
    for s in range(base):
        x = rng.uniform(-1.0, 1.0, vals).astype(np.float32)
        if s in (0, 17, 42, 63):
            keep[s] = x
        d_in.upload(x, s * vals)
    for g in range(1, n // base):      # device-to-device copies of the 64 base streams
        assert lib.rsb_memcpy(0, d_in.ptr + 4 * g * base * vals, d_in.ptr, base * vals * 4, 2) == 0
    out_stride = ((int(frames * 48000 / 44100) + 8) * ch + 3) & ~3
    d_out, (cons, prod, calls) = run_device(Kernel.TENSOR, d_in, n, ch, frames, 44100, 48000, 3, 512, out_stride)
    subset = {0: 0, 17 + 64: 17, 42 + 7 * 64: 42, 63 + 15 * 64: 63, 17 + 9 * 64: 17}   # stream -> base stream
    refs = {b: O.OracleFir(ch, 44100, 48000, 3, 1).process(keep[b], 512 * ch) for b in set(subset.values())}
    r0 = refs[0]
    assert all(c == r0["consumed_total"] for c in cons) and all(p == len(r0["out"]) for p in prod)
    assert all(k == r0["calls"] for k in calls) and calls[0] == 5168 and prod[0] == 2879862 * ch
    worst, total = 0.0, 0
    for s, b in subset.items():
        got = d_out.download(prod[s], s * out_stride)
        d = np.abs(got.astype(np.float64) - refs[b]["out"])
        worst = max(worst, float(d.max()))
        total += d.size
    print(f"\nfull-size configs[1], uniform noise: {total:.3e} samples of 5 streams vs oracle, max |diff| {worst:.3e}")
    assert worst <= TOL_MARGIN, worst
    d_in.free()
    d_out.free()


def test_non_finite_and_out_of_range_samples_contract():
    """include/resampler_b200.h, "non-finite samples": EXACT follows IEEE like the reference
    (same outputs become Inf / NaN, bit-identical elsewhere); TENSOR / FAST leave every output
    whose tile does not read the bad sample untouched (other streams, the other channel, outputs
    more than a tile + window away) and make no promise for the rest."""
    n, ch, frames = 70, 2, 6000
    rng = np.random.default_rng(9)
    xs = [rng.uniform(-1.0, 1.0, frames * ch).astype(np.float32) for _ in range(n)]
    clean = [x.copy() for x in xs]
    bad = {3: (2500, 0, np.float32(np.inf)), 5: (1000, 1, np.float32(np.nan)), 7: (4000, 0, np.float32(1.0e5))}
    for s, (f, c, v) in bad.items():
        xs[s][f * ch + c] = v
    e = FirBatch(n, ch, 44100, 48000, Latency.Sample64, Attenuation.Db90, kernel=Kernel.EXACT)
    got_e = [np.array(o, copy=True) for o in e.process(xs, 512 * ch)["out"]]
    e.close()
    e2 = FirBatch(n, ch, 44100, 48000, Latency.Sample64, Attenuation.Db90, kernel=Kernel.EXACT)
    ref_clean = [np.array(o, copy=True) for o in e2.process(clean, 512 * ch)["out"]]
    e2.close()
    for s in bad:      # EXACT == oracle: same non-finite outputs, bit-identical finite ones
        ref = O.OracleFir(ch, 44100, 48000, 3, 1).process(xs[s], 512 * ch)["out"]
        assert np.array_equal(np.isnan(got_e[s]), np.isnan(ref)), s
        fin = ~np.isnan(ref)
        assert np.array_equal(bits(got_e[s][fin]), bits(ref[fin])), s
    ratio = 44100 / 48000
    for kern, tile in ((Kernel.TENSOR, 64), (Kernel.FAST, 32)):
        b = FirBatch(n, ch, 44100, 48000, Latency.Sample64, Attenuation.Db90, kernel=kern)
        got = b.process(xs, 512 * ch)["out"]
        assert b.last_kernel() == kern
        for s in range(n):
            g = got[s].astype(np.float64).reshape(-1, ch)
            want = ref_clean[s].astype(np.float64).reshape(-1, ch)
            if s not in bad:
                assert np.max(np.abs(g - want)) <= TOL, (kern, s)
                continue
            f, c, _ = bad[s]
            # outputs whose window or tile can touch input frame f: generous bracket in output frames
            lo = int((f - 128 - 32) / ratio) - 2 * tile
            hi = int((f + 32) / ratio) + 2 * tile
            far = np.ones(len(g), bool)
            far[max(lo, 0):hi] = False
            assert np.max(np.abs(g[far] - want[far])) <= TOL, (kern, s, "far outputs")
            assert np.max(np.abs(g[:, 1 - c] - want[:, 1 - c])) <= TOL, (kern, s, "other channel")
        b.close()
