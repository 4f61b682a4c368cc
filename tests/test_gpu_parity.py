"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle.

Bar (BASELINE.json north_star): (consumed, produced) counts and phase indices bit-exact;
samples bit-identical for the EXACT kernel (the reference's AVX-512 summation order) and
within 1e-6 absolute for the FAST kernel.
"""
import numpy as np
import pytest

import oracle_lib as O
from resampler_b200 import (Attenuation, FirBatch, Kernel, Latency, ResampleError, ResamplerFir,
                            SampleRate)
from resampler_b200.fir import FLAG_KEEP_PLAN, FLAG_RECORD_CALLS

pytestmark = pytest.mark.gpu

TOL_FAST = 1e-6   # absolute, north_star


def noise(rng, n):
    return rng.uniform(-1.0, 1.0, n).astype(np.float32)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def oracle_stream(ch, in_hz, out_hz, lat, att, x, call_len, out_cap_len=0, trace=False):
    f = O.OracleFir(ch, in_hz, out_hz, lat, att)
    return f.process(x, call_len, out_cap_len=out_cap_len, trace=trace)


# ---------------------------------------------------------------------------------------------
# single-stream mirror of the reference API
# ---------------------------------------------------------------------------------------------
CONFIGS = [
    # ch, in_hz, out_hz, latency, attenuation, call frames
    (2, 48000, 44100, 3, 1, 512),    # BASELINE config 1
    (2, 44100, 48000, 3, 1, 512),    # config 2 parameters
    (1, 16000, 48000, 1, 1, 160),    # config 3 parameters
    (8, 96000, 48000, 2, 1, 512),    # config 4 parameters
    (1, 22050, 44100, 3, 1, 256),    # the reference's stop-band test rates
    (3, 44100, 48000, 0, 0, 100),    # odd channel count, 16 taps, Db60
    (2, 24000, 16000, 2, 2, 333),    # arbitrary rates (new_from_hz), Db120
]


@pytest.mark.parametrize("ch,in_hz,out_hz,lat,att,call", CONFIGS)
def test_streaming_resample_bit_exact(ch, in_hz, out_hz, lat, att, call):
    """The canonical caller loop over ResamplerFir.resample(): counts and samples bit-exact."""
    rng = np.random.default_rng(ch * 1000 + call)
    frames = call * 12 + 37
    x = noise(rng, frames * ch)
    ref = O.OracleFir(ch, in_hz, out_hz, lat, att)
    dut = ResamplerFir.new_from_hz(ch, in_hz, out_hz, Latency(lat), Attenuation(att),
                                   kernel=Kernel.EXACT)
    assert dut.buffer_size_output() == ref.buffer_size_output()
    assert dut.delay() == ref.delay()
    out_ref = np.zeros(ref.buffer_size_output(), np.float32)
    out_dut = np.zeros(dut.buffer_size_output(), np.float32)
    off = 0
    while off < len(x):
        chunk = x[off:off + call * ch]
        err, c, p = ref.resample(chunk, out_ref)
        c2, p2 = dut.resample(chunk, out_dut)
        assert err == 0 and (c2, p2) == (c, p)
        assert np.array_equal(bits(out_dut[:p]), bits(out_ref[:p]))
        off += c
        if c == 0:
            break
    dut.close()


def test_new_equals_new_from_hz_and_reset():
    """src/resampler_fir.rs:817-839 plus reset() (:638-642)."""
    a = ResamplerFir(1, SampleRate.Hz48000, SampleRate.Hz44100, Latency.Sample64, Attenuation.Db90,
                     kernel=Kernel.EXACT)
    b = ResamplerFir.new_from_hz(1, 48000, 44100, Latency.Sample64, Attenuation.Db90,
                                 kernel=Kernel.EXACT)
    x = np.full(512, 0.5, np.float32)
    oa = np.zeros(a.buffer_size_output(), np.float32)
    ob = np.zeros(b.buffer_size_output(), np.float32)
    ra, rb = a.resample(x, oa), b.resample(x, ob)
    assert ra == rb and np.array_equal(bits(oa[:ra[1]]), bits(ob[:rb[1]]))
    ref = O.OracleFir(1, 48000, 44100, 3, 1)
    oref = np.zeros(ref.buffer_size_output(), np.float32)
    _, c, p = ref.resample(x, oref)
    assert ra == (c, p) and np.array_equal(bits(oa[:p]), bits(oref[:p]))
    # reset: the next call behaves like the first one
    a.reset()
    oa2 = np.zeros_like(oa)
    assert a.resample(x, oa2) == ra and np.array_equal(bits(oa2[:ra[1]]), bits(oa[:ra[1]]))
    a.close()
    b.close()


def test_arbitrary_rates_and_errors():
    """src/resampler_fir.rs:841-862 and the error precedence of :514-519."""
    r = ResamplerFir.new_from_hz(1, 24000, 16000, Latency.Sample32, Attenuation.Db60)
    out = np.zeros(r.buffer_size_output(), np.float32)
    r.resample(np.zeros(256, np.float32), out)
    r.close()
    with pytest.raises(ValueError, match="input sample rate must be greater than zero"):
        ResamplerFir.new_from_hz(1, 0, 44100)
    with pytest.raises(ValueError, match="output sample rate must be greater than zero"):
        ResamplerFir.new_from_hz(1, 44100, 0)
    s = ResamplerFir.new_from_hz(2, 48000, 44100)
    with pytest.raises(ResampleError) as e:
        s.resample(np.zeros(3, np.float32), np.zeros(3, np.float32))
    assert e.value.kind == ResampleError.InvalidInputBufferSize
    with pytest.raises(ResampleError) as e:
        s.resample(np.zeros(4, np.float32), np.zeros(3, np.float32))
    assert e.value.kind == ResampleError.InvalidOutputBufferSize
    # a failed call must not have touched the state
    ref = O.OracleFir(2, 48000, 44100, 3, 2)
    x = noise(np.random.default_rng(3), 1024)
    o1 = np.zeros(s.buffer_size_output(), np.float32)
    o2 = np.zeros(ref.buffer_size_output(), np.float32)
    c, p = s.resample(x, o1)
    _, c2, p2 = ref.resample(x, o2)
    assert (c, p) == (c2, p2) and np.array_equal(bits(o1[:p]), bits(o2[:p]))
    s.close()


def test_edge_cases_empty_small_capacity_and_lookahead():
    """Empty input, output capacity smaller than what is available (:553, :528), a full
    4096-frame buffer, ratio > taps (the 0.5.1 fix, :593-596)."""
    rng = np.random.default_rng(11)
    for (in_hz, out_hz, lat, calls) in [
            (44100, 48000, 3, [(0, 8642), (512, 100), (512, 100), (0, 100), (4096, 50), (4096, 0),
                               (9000, 8642), (1, 1), (0, 8642), (4096, 8642), (4096, 8642)]),
            (384000, 1000, 0, [(512, 100)] * 8 + [(0, 5), (4096, 3)]),
            (8000, 192000, 3, [(300, 1000), (300, 98306), (10, 98306), (0, 98306)])]:
        ref = O.OracleFir(1, in_hz, out_hz, lat, 1)
        dut = ResamplerFir.new_from_hz(1, in_hz, out_hz, Latency(lat), Attenuation.Db90,
                                       kernel=Kernel.EXACT)
        for n_in, cap in calls:
            x = noise(rng, n_in)
            o_ref = np.zeros(cap, np.float32)
            o_dut = np.zeros(cap, np.float32)
            _, c, p = ref.resample(x, o_ref)
            assert dut.resample(x, o_dut) == (c, p), (n_in, cap)
            assert np.array_equal(bits(o_dut[:p]), bits(o_ref[:p]))
        dut.close()


# ---------------------------------------------------------------------------------------------
# batched submit: n independent resample() calls
# ---------------------------------------------------------------------------------------------
def test_submit_batch_divergent_streams_bit_exact():
    """BASELINE config 3 (ii): mono 16->48 kHz, 32 taps, every stream with its own pseudo-random
    call sizes => divergent state; counts, phases and samples bit-exact per stream."""
    n = 24
    rng = np.random.default_rng(5)
    sizes = [1, 7, 16, 33, 160, 480]
    batch = FirBatch(n, 1, 16000, 48000, Latency.Sample16, Attenuation.Db90, kernel=Kernel.EXACT)
    refs = [O.OracleFir(1, 16000, 48000, 1, 1) for _ in range(n)]
    bso = batch.buffer_size_output()
    for it in range(14):
        ins = [noise(rng, int(rng.choice(sizes))) for _ in range(n)]
        outs = [np.zeros(bso, np.float32) for _ in range(n)]
        cons, prod = batch.submit(ins, outs, flags=FLAG_KEEP_PLAN)
        for s in range(n):
            o = np.zeros(bso, np.float32)
            _, c, p, tr = refs[s].resample(ins[s], o, trace=True)
            assert (cons[s], prod[s]) == (c, p), (it, s)
            assert np.array_equal(bits(outs[s][:p]), bits(o[:p])), (it, s)
            if s % 5 == 0:
                plan = batch.last_plan(s)
                for key in ("input_offset", "phase1", "phase2", "frac_bits"):
                    assert np.array_equal(plan[key], tr[key]), (key, it, s)
    batch.close()


def test_submit_batch_subset_of_streams_and_mixed_capacities():
    n = 6
    rng = np.random.default_rng(8)
    batch = FirBatch(n, 2, 44100, 48000, Latency.Sample64, Attenuation.Db90, kernel=Kernel.EXACT)
    refs = [O.OracleFir(2, 44100, 48000, 3, 1) for _ in range(n)]
    for it in range(6):
        streams = [int(s) for s in rng.permutation(n)[:4]]
        ins = [noise(rng, 2 * int(rng.integers(0, 900))) for _ in streams]
        caps = [2 * int(rng.choice([0, 10, 700, 4321])) for _ in streams]
        outs = [np.zeros(c, np.float32) for c in caps]
        cons, prod = batch.submit(ins, outs, streams=streams)
        for j, s in enumerate(streams):
            o = np.zeros(caps[j], np.float32)
            _, c, p = refs[s].resample(ins[j], o)
            assert (cons[j], prod[j]) == (c, p)
            assert np.array_equal(bits(outs[j][:p]), bits(o[:p]))
    with pytest.raises(ValueError):
        batch.submit([np.zeros(2, np.float32)] * 2, [np.zeros(2, np.float32)] * 2, streams=[1, 1])
    batch.close()


# ---------------------------------------------------------------------------------------------
# batched multi-call: the canonical caller loop per stream in one launch
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ch,in_hz,out_hz,lat,call_frames,cap_frames,n_streams", [
    (2, 44100, 48000, 3, 512, 0, 5),      # config 2 pattern
    (2, 48000, 44100, 3, 512, 0, 3),      # config 1 pattern
    (1, 16000, 48000, 1, 160, 0, 7),      # config 3 (i)
    (8, 96000, 48000, 2, 512, 0, 3),      # config 4
    (2, 44100, 48000, 3, 512, 100, 2),    # capacity-limited calls
    (1, 384000, 1000, 0, 512, 0, 2),      # ratio > taps
    (2, 44100, 48000, 3, 4096, 0, 2),     # maximum call size
])
def test_process_batch_bit_exact(ch, in_hz, out_hz, lat, call_frames, cap_frames, n_streams):
    rng = np.random.default_rng(in_hz + call_frames)
    frames = min(in_hz // 3, 20000) + 13
    xs = [noise(rng, frames * ch) for _ in range(n_streams)]
    batch = FirBatch(n_streams, ch, in_hz, out_hz, Latency(lat), Attenuation.Db90,
                     kernel=Kernel.EXACT)
    res = batch.process(xs, call_frames * ch, cap_frames * ch,
                        flags=FLAG_RECORD_CALLS | FLAG_KEEP_PLAN)
    for s in range(n_streams):
        ref = oracle_stream(ch, in_hz, out_hz, lat, 1, xs[s], call_frames * ch, cap_frames * ch,
                            trace=(s == 0))
        assert res["consumed"][s] == ref["consumed_total"]
        assert res["produced"][s] == len(ref["out"])
        assert res["calls"][s] == ref["calls"]
        assert np.array_equal(bits(res["out"][s]), bits(ref["out"])), s
        if s == 0:
            cc, cp = batch.last_call_counts(0)
            assert np.array_equal(cc, ref["consumed"]) and np.array_equal(cp, ref["produced"])
            plan = batch.last_plan(0)
            for key in ("input_offset", "phase1", "phase2", "frac_bits"):
                assert np.array_equal(plan[key], ref["trace"][key]), key
    # a second batch continues from the carried state (history + phase), like more calls
    xs2 = [noise(rng, 777 * ch) for _ in range(n_streams)]
    res2 = batch.process(xs2, call_frames * ch, cap_frames * ch)
    for s in range(n_streams):
        f = O.OracleFir(ch, in_hz, out_hz, lat, 1)
        f.process(xs[s], call_frames * ch, out_cap_len=cap_frames * ch)
        ref2 = f.process(xs2[s], call_frames * ch, out_cap_len=cap_frames * ch)
        assert res2["produced"][s] == len(ref2["out"])
        assert np.array_equal(bits(res2["out"][s]), bits(ref2["out"])), s
    batch.close()


def test_process_batch_ragged_lengths_and_empty():
    """Streams of different lengths (incl. empty) in one batch => several plan units."""
    ch, n = 2, 6
    rng = np.random.default_rng(21)
    lens = [0, 1, 127, 128, 5000, 12345]
    xs = [noise(rng, L * ch) for L in lens]
    batch = FirBatch(n, ch, 44100, 48000, Latency.Sample64, Attenuation.Db90, kernel=Kernel.EXACT)
    res = batch.process(xs, 512 * ch)
    for s in range(n):
        ref = oracle_stream(ch, 44100, 48000, 3, 1, xs[s], 512 * ch)
        assert res["consumed"][s] == ref["consumed_total"] and res["calls"][s] == ref["calls"]
        assert np.array_equal(bits(res["out"][s]), bits(ref["out"]))
    batch.close()


def test_process_batch_output_capacity_error():
    batch = FirBatch(1, 1, 44100, 48000, Latency.Sample64, Attenuation.Db90, kernel=Kernel.EXACT)
    x = np.zeros(5000, np.float32)
    with pytest.raises(RuntimeError, match="too small"):
        batch.process([x], 512, out_capacity=100)
    batch.close()


def test_stopband_attenuation_on_gpu():
    """src/resampler_fir.rs:741-815 replayed through the CUDA path (numpy FFT for analysis)."""
    for in_hz, out_hz in [(22050, 44100), (22050, 48000)]:
        n = int(np.float32(in_hz) * np.float32(5.0))
        x = np.zeros(n, np.float32)
        x[n // 2] = 1.0
        b = FirBatch(1, 1, in_hz, out_hz, Latency.Sample64, Attenuation.Db90)
        y = b.process([x], 256)["out"][0]
        b.close()
        peak = int(np.argmax(np.abs(y)))
        win = int(out_hz * 0.1)
        start = max(peak - win // 2, 0)
        ir = y[start:min(start + win, len(y))]
        buf = np.zeros(8192)
        m = min(len(ir), 8192)
        buf[:m] = ir[:m]
        mag = np.abs(np.fft.rfft(buf))
        db = np.where(mag > 1e-10, 20 * np.log10(np.maximum(mag, 1e-300)), -200.0)
        tb = lambda f: int(round(f / out_hz * 8192))  # noqa: E731
        nyq = in_hz / 2
        pb = db[tb(20.0):tb(nyq * 0.9) + 1].max()
        sb = db[tb(nyq * 1.1):min(len(db) - 10, tb(out_hz / 2 * 0.95)) + 1].max()
        assert pb - sb >= 90.0


# ---------------------------------------------------------------------------------------------
# FAST kernel: counts/phases bit-exact, samples within 1e-6 absolute (north_star tolerance)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ch,in_hz,out_hz,lat,call_frames,cap_frames,n_streams", [
    (2, 44100, 48000, 3, 512, 0, 70),     # config 2 pattern; 70 streams = one full + one partial group
    (2, 48000, 44100, 3, 512, 0, 9),      # config 1 pattern
    (1, 16000, 48000, 1, 160, 0, 130),    # config 3 (i), mono, 32 taps
    (8, 96000, 48000, 2, 512, 0, 17),     # config 4, 8 channels, 64 taps
    (3, 44100, 48000, 0, 100, 0, 5),      # odd channel count, 16 taps
    (2, 44100, 48000, 3, 512, 100, 4),    # capacity-limited calls
    (2, 22050, 48000, 3, 4096, 0, 3),     # maximum call size
    (1, 192000, 48000, 3, 512, 0, 40),    # ratio 4
])
def test_fast_kernel_within_tolerance(ch, in_hz, out_hz, lat, call_frames, cap_frames, n_streams):
    check_kernel_within_tolerance(Kernel.FAST, ch, in_hz, out_hz, lat, call_frames, cap_frames,
                                  n_streams)


# ---------------------------------------------------------------------------------------------
# TENSOR kernel (tcgen05, split fp16 with prescale): same bar -- counts/phases bit-exact, samples within 1e-6
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ch,in_hz,out_hz,lat,call_frames,cap_frames,n_streams", [
    (2, 44100, 48000, 3, 512, 0, 70),     # config 2 pattern; one full + one partial row group
    (2, 48000, 44100, 3, 512, 0, 9),      # config 1 pattern
    (1, 16000, 48000, 1, 160, 0, 130),    # config 3 (i), mono, 32 taps; 128 + 2 streams
    (1, 44100, 48000, 3, 512, 0, 200),    # mono, 128 taps
    (2, 44100, 48000, 0, 100, 0, 5),      # 16 taps
    (2, 44100, 48000, 3, 512, 100, 4),    # capacity-limited calls
    (2, 22050, 48000, 3, 4096, 0, 3),     # maximum call size
    (1, 96000, 48000, 2, 512, 0, 40),     # ratio 2, 64 taps
    (8, 96000, 48000, 2, 512, 0, 17),     # config 4: 8 channels, 64 taps; 16 + 1 members
    (4, 44100, 48000, 3, 512, 0, 40),     # 4 channels, 128 taps
])
def test_tensor_kernel_within_tolerance(ch, in_hz, out_hz, lat, call_frames, cap_frames, n_streams):
    check_kernel_within_tolerance(Kernel.TENSOR, ch, in_hz, out_hz, lat, call_frames, cap_frames,
                                  n_streams)


def test_tensor_kernel_falls_back_for_ragged_batches_and_wide_ratios():
    """Streams with different sizes do not share a plan: the batch runs on FAST / EXACT and is
    still right; a ratio whose tile window does not fit the TMEM ring is refused up front."""
    n, ch = 6, 2
    rng = np.random.default_rng(5)
    batch = FirBatch(n, ch, 44100, 48000, Latency.Sample64, Attenuation.Db90, kernel=Kernel.TENSOR)
    xs = [noise(rng, (3000 + 37 * s) * ch) for s in range(n)]
    res = batch.process(xs, 512 * ch, 0)
    assert batch.last_kernel() != Kernel.TENSOR
    for s in range(n):
        ref = oracle_stream(ch, 44100, 48000, 3, 1, xs[s], 512 * ch, 0)
        assert res["produced"][s] == len(ref["out"])
        assert np.max(np.abs(res["out"][s].astype(np.float64) - ref["out"])) <= TOL_FAST
    batch.close()
    with pytest.raises(ValueError):
        FirBatch(4, 1, 192000, 48000, Latency.Sample64, Attenuation.Db90, kernel=Kernel.TENSOR)


def test_tensor_kernel_short_tiles_many_items():
    """Regression: mono 16 -> 48 kHz with 32 taps gives very short tiles (7 K steps), so the two
    MMA issuers, the G producer and the epilogue interleave tightly, and every CTA works through
    several items.  An mbarrier parity wait placed before its gating wait used to pass early here
    (illegal instruction in ~30 % of the runs).  TENSOR against EXACT on every stream, 3 times."""
    n, ch, frames = 512, 1, 64000
    rng = np.random.default_rng(11)
    xs = [noise(rng, frames) for _ in range(n)]
    ref = FirBatch(n, ch, 16000, 48000, Latency.Sample16, Attenuation.Db90, kernel=Kernel.EXACT)
    want = [np.array(o, copy=True) for o in ref.process(xs, 160, 0)["out"]]
    ref.close()
    o0 = oracle_stream(ch, 16000, 48000, 1, 1, xs[0], 160, 0)["out"]
    assert np.array_equal(bits(want[0]), bits(o0))
    for _ in range(3):
        b = FirBatch(n, ch, 16000, 48000, Latency.Sample16, Attenuation.Db90, kernel=Kernel.TENSOR)
        got = b.process(xs, 160, 0)["out"]
        assert b.last_kernel() == Kernel.TENSOR
        for s in range(n):
            assert len(got[s]) == len(want[s])
            assert np.max(np.abs(got[s].astype(np.float64) - want[s])) <= TOL_FAST, s
        b.close()


def check_kernel_within_tolerance(kernel, ch, in_hz, out_hz, lat, call_frames, cap_frames,
                                  n_streams):
    rng = np.random.default_rng(in_hz + call_frames + n_streams)
    frames = min(in_hz // 4, 9000) + 29
    xs = [noise(rng, frames * ch) for _ in range(n_streams)]
    batch = FirBatch(n_streams, ch, in_hz, out_hz, Latency(lat), Attenuation.Db90, kernel=kernel)
    res = batch.process(xs, call_frames * ch, cap_frames * ch, flags=FLAG_KEEP_PLAN)
    assert batch.last_kernel() == kernel
    worst = 0.0
    for s in range(n_streams):
        ref = oracle_stream(ch, in_hz, out_hz, lat, 1, xs[s], call_frames * ch, cap_frames * ch,
                            trace=(s == 0))
        assert res["consumed"][s] == ref["consumed_total"]
        assert res["produced"][s] == len(ref["out"])
        assert res["calls"][s] == ref["calls"]
        d = np.abs(res["out"][s].astype(np.float64) - ref["out"].astype(np.float64))
        worst = max(worst, float(d.max()) if d.size else 0.0)
        if s == 0:
            plan = batch.last_plan(0)
            for key in ("input_offset", "phase1", "phase2", "frac_bits"):
                assert np.array_equal(plan[key], ref["trace"][key]), key
    assert worst <= TOL_FAST, worst
    # state carry into a second batch (history written back by the fast path's update)
    xs2 = [noise(rng, 601 * ch) for _ in range(n_streams)]
    res2 = batch.process(xs2, call_frames * ch, cap_frames * ch)
    for s in range(0, n_streams, max(1, n_streams // 5)):
        f = O.OracleFir(ch, in_hz, out_hz, lat, 1)
        f.process(xs[s], call_frames * ch, out_cap_len=cap_frames * ch)
        ref2 = f.process(xs2[s], call_frames * ch, out_cap_len=cap_frames * ch)
        assert res2["produced"][s] == len(ref2["out"])
        d = np.abs(res2["out"][s].astype(np.float64) - ref2["out"].astype(np.float64))
        assert (float(d.max()) if d.size else 0.0) <= TOL_FAST
    batch.close()


def test_fast_kernel_streaming_submit_and_divergence():
    """Small-chunk streaming through submit(): first a shared cohort, then divergent sizes."""
    n, ch = 40, 2
    rng = np.random.default_rng(77)
    batch = FirBatch(n, ch, 44100, 48000, Latency.Sample64, Attenuation.Db90, kernel=Kernel.FAST)
    refs = [O.OracleFir(ch, 44100, 48000, 3, 1) for _ in range(n)]
    bso = batch.buffer_size_output()
    for it in range(10):
        if it < 5:
            sizes = [512] * n
        else:
            sizes = [int(rng.choice([0, 1, 64, 300, 512])) for _ in range(n)]
        ins = [noise(rng, sz * ch) for sz in sizes]
        outs = [np.zeros(bso, np.float32) for _ in range(n)]
        cons, prod = batch.submit(ins, outs)
        for s in range(n):
            o = np.zeros(bso, np.float32)
            _, c, p = refs[s].resample(ins[s], o)
            assert (cons[s], prod[s]) == (c, p), (it, s)
            if p:
                assert np.max(np.abs(outs[s][:p].astype(np.float64) - o[:p])) <= TOL_FAST
    batch.close()


def test_device_memspace_async_and_full_size_properties():
    """Device-resident buffers, async submit; at a larger size checks size-independent
    properties: linearity (2x input -> 2x output exactly, scaling by 2 is exact in fp32) and
    determinism, for every kernel."""
    from resampler_b200.fir import FLAG_ASYNC, MEM_DEVICE, DeviceBuffer
    from resampler_b200 import _lib
    n, ch, frames = 256, 2, 44100
    lib = _lib.load()
    d_in = DeviceBuffer(0, n * frames * ch)
    assert lib.rsb_fill_synthetic(0, d_in.ptr, 0, n, frames, ch, 44100, 0x5EED) == 0
    host = d_in.download().reshape(n, frames * ch)
    d_in2 = DeviceBuffer(0, n * frames * ch)
    d_in2.upload((host * np.float32(2.0)).astype(np.float32))
    out_stride = (int(frames * 48000 / 44100) + 8) * ch
    n_expected = len(oracle_stream(ch, 44100, 48000, 3, 1, host[0], 512 * ch)["out"])
    outs = {}
    for kern in (Kernel.EXACT, Kernel.FAST, Kernel.TENSOR):
        b = FirBatch(n, ch, 44100, 48000, Latency.Sample64, Attenuation.Db90, kernel=kern)
        for name, src in (("x", d_in), ("2x", d_in2), ("x_again", d_in)):
            b.reset(-1)
            d_out = DeviceBuffer(0, n * out_stride)
            cons, prod, calls = b.process_ptrs(
                [src.ptr + 4 * s * frames * ch for s in range(n)], [frames * ch] * n, 512 * ch, 0,
                [d_out.ptr + 4 * s * out_stride for s in range(n)], [out_stride] * n,
                memspace=MEM_DEVICE, flags=FLAG_ASYNC)
            b.sync()
            assert all(c == frames * ch for c in cons[:])
            assert len(set(prod[:])) == 1 and prod[0] == n_expected
            outs[(kern, name)] = d_out.download().reshape(n, out_stride)[:, :prod[0]]
            d_out.free()
        assert np.array_equal(outs[(kern, "x")], outs[(kern, "x_again")])
        if kern == Kernel.TENSOR:
            # fp16 hi/lo operands: the lo part of a sample below 2^-6 is an fp16 subnormal (absolute
            # quantum 2^-28 in units of x); that ~1e-9 perturbation can flip the last bit of an
            # accumulation, so scaling by two is exact to ONE ulp of the output (|y| < 4)
            dev = float(np.max(np.abs(outs[(kern, "2x")].astype(np.float64) -
                                      2.0 * outs[(kern, "x")].astype(np.float64))))
            assert dev <= 2.4e-7, dev
        else:
            assert np.array_equal(outs[(kern, "2x")], outs[(kern, "x")] * np.float32(2.0))
        b.close()
    # the kernels agree within the tolerance on every sample of every stream
    for kern in (Kernel.FAST, Kernel.TENSOR):
        assert np.max(np.abs(outs[(kern, "x")].astype(np.float64) -
                             outs[(Kernel.EXACT, "x")])) <= TOL_FAST, kern
    # and stream 0 / stream n-1 equal the oracle (EXACT: bit for bit)
    for s in (0, n - 1):
        ref = oracle_stream(ch, 44100, 48000, 3, 1, host[s], 512 * ch)
        assert np.array_equal(bits(outs[(Kernel.EXACT, "x")][s]), bits(ref["out"]))


def test_flush_feeds_delay_frames_of_silence():
    """Opt-in tail handling (not in the reference): flush == resample(zeros(delay * ch)),
    checked against the oracle doing exactly that; the state carries on afterwards."""
    ch, n = 2, 3
    rng = np.random.default_rng(33)
    xs = [noise(rng, L * ch) for L in (3000, 2500, 64)]
    batch = FirBatch(n, ch, 44100, 48000, Latency.Sample64, Attenuation.Db90, kernel=Kernel.EXACT)
    batch.process(xs, 512 * ch)
    tails = batch.flush()
    more = batch.process([xs[0][:400]] * n, 512 * ch)
    for s in range(n):
        f = O.OracleFir(ch, 44100, 48000, 3, 1)
        f.process(xs[s], 512 * ch)
        z = np.zeros(f.delay() * ch, np.float32)
        ref = f.process(z, len(z))
        assert len(tails[s]) == len(ref["out"]) and len(tails[s]) > 0
        assert np.array_equal(bits(tails[s]), bits(ref["out"])), s
        ref2 = f.process(xs[0][:400], 512 * ch)
        assert np.array_equal(bits(more["out"][s]), bits(ref2["out"])), s
    # a subset of the streams
    t1 = batch.flush(streams=[1])
    assert len(t1) == 1
    batch.close()


def test_cuda_path_matches_committed_fixtures():
    """The CUDA path (bit-exact kernel, through the C ABI) against the frozen fixtures of
    tests/golden/fir_cases.npz: per-call counts, per-frame plan and output sample bits."""
    import importlib.util
    from pathlib import Path
    gdir = Path(__file__).resolve().parent / "golden"
    spec = importlib.util.spec_from_file_location("make_fixtures", gdir / "make_fixtures.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    z = np.load(gdir / "fir_cases.npz")
    for name, (ch, in_hz, out_hz, lat, att, call, cap, frames) in mod.CASES.items():
        batch = FirBatch(1, ch, in_hz, out_hz, Latency(lat), Attenuation(att), kernel=Kernel.EXACT)
        res = batch.process([z[f"{name}/x"]], call * ch, cap * ch,
                            flags=FLAG_RECORD_CALLS | FLAG_KEEP_PLAN)
        assert np.array_equal(bits(res["out"][0]), bits(z[f"{name}/out"])), name
        cc, cp = batch.last_call_counts(0)
        assert np.array_equal(cc, z[f"{name}/consumed"]) and np.array_equal(cp, z[f"{name}/produced"])
        plan = batch.last_plan(0)
        for key in ("input_offset", "phase1", "phase2", "frac_bits"):
            assert np.array_equal(plan[key], z[f"{name}/{key}"]), (name, key)
        batch.close()


def test_fused_submit_divergent_streams_bit_exact_and_handoff():
    """The single-launch submit path (fir_submit.cu, device memspace): per-stream pseudo-random call
    sizes (BASELINE configs[2] variant (ii)) -> every stream has its own plan.  Counts and samples
    bit-identical to the oracle on every call; afterwards a big batch on the same handle (the tile
    kernels, host-side plan from the refreshed mirror) carries on from the right state."""
    from resampler_b200.fir import MEM_DEVICE, DeviceBuffer
    n, ch, in_hz, out_hz, lat = 96, 1, 16000, 48000, 1
    sizes = [0, 1, 7, 16, 33, 160, 480]
    rng = np.random.default_rng(2)
    batch = FirBatch(n, ch, in_hz, out_hz, Latency(lat), Attenuation.Db90, kernel=Kernel.AUTO)
    refs = [O.OracleFir(ch, in_hz, out_hz, lat, 1) for _ in range(n)]
    bso = batch.buffer_size_output()
    d_in = DeviceBuffer(0, n * 480 * ch)
    d_out = DeviceBuffer(0, n * bso)
    for it in range(25):
        if it < 3:
            sz = [160] * n                      # a shared cohort first
        else:
            sz = [int(rng.choice(sizes)) for _ in range(n)]
        ins = [noise(rng, s * ch) for s in sz]
        for s in range(n):
            if sz[s]:
                d_in.upload(ins[s], s * 480 * ch)
        cons, prod = batch.submit_ptrs([d_in.ptr + 4 * s * 480 * ch for s in range(n)], [x.size for x in ins],
                                       [d_out.ptr + 4 * s * bso for s in range(n)], [bso] * n,
                                       memspace=MEM_DEVICE)
        assert batch.last_kernel() == Kernel.EXACT
        for s in range(n):
            o = np.zeros(bso, np.float32)
            _, c, p = refs[s].resample(ins[s], o)
            assert (cons[s], prod[s]) == (c, p), (it, s)
            if p:
                assert np.array_equal(bits(d_out.download(p, s * bso)), bits(o[:p])), (it, s)
    # hand-off to the batch path: the same continuation for every stream, 3000 frames
    tail = [noise(rng, 3000 * ch) for _ in range(n)]
    res = batch.process(tail, 160 * ch)
    for s in range(0, n, 7):
        ref = refs[s].process(tail[s], 160 * ch)
        assert res["produced"][s] == len(ref["out"]), s
        assert np.max(np.abs(res["out"][s].astype(np.float64) - ref["out"])) <= TOL_FAST, s
    batch.close()
    d_in.free()
    d_out.free()


def test_host_pipeline_equals_unsliced_batch(monkeypatch):
    """Host-memspace process() of a long batch runs as a pipeline of time slices (H2D, kernels and
    D2H overlapped, slice boundaries from a dry run of the plan).  Counts, call numbers and -- with
    the bit-exact kernel -- every sample equal the unsliced path's; the state carries on correctly
    afterwards; the tensor kernel's sliced result stays within the bar of the oracle."""
    n, ch, in_hz, out_hz, lat = 48, 2, 44100, 48000, 3
    frames = 200_001                      # not a multiple of the call size
    rng = np.random.default_rng(11)
    ins = [noise(rng, frames * ch) for _ in range(n)]
    tail = [noise(rng, 3000 * ch) for _ in range(n)]
    results = {}
    for mode in ("pipelined", "unsliced"):
        if mode == "unsliced":
            monkeypatch.setenv("RSB_NO_HOST_PIPELINE", "1")
        else:
            monkeypatch.delenv("RSB_NO_HOST_PIPELINE", raising=False)
        batch = FirBatch(n, ch, in_hz, out_hz, Latency(lat), Attenuation.Db90, kernel=Kernel.EXACT)
        r1 = batch.process(ins, 512 * ch)
        b, sl = batch.host_pipeline_stats()
        assert (b, sl > 2) == ((1, True) if mode == "pipelined" else (0, False))
        r2 = batch.process(tail, 160 * ch)          # short: always unsliced, carries on from the state
        results[mode] = (r1, r2)
        batch.close()
    (p1, p2), (u1, u2) = results["pipelined"], results["unsliced"]
    for a, b_ in ((p1, u1), (p2, u2)):
        assert np.array_equal(a["consumed"], b_["consumed"])
        assert np.array_equal(a["produced"], b_["produced"])
        assert np.array_equal(a["calls"], b_["calls"])
        for s in range(n):
            assert np.array_equal(bits(a["out"][s]), bits(b_["out"][s])), s
    # against the oracle (two streams), then the tensor kernel through the same pipeline
    monkeypatch.delenv("RSB_NO_HOST_PIPELINE", raising=False)
    refs = {}
    for s in (0, n - 1):
        ref = O.OracleFir(ch, in_hz, out_hz, lat, 1)
        refs[s] = ref.process(ins[s], 512 * ch)
        assert p1["produced"][s] == len(refs[s]["out"])
        assert np.array_equal(bits(p1["out"][s]), bits(np.asarray(refs[s]["out"], np.float32)))
    batch = FirBatch(n, ch, in_hz, out_hz, Latency(lat), Attenuation.Db90, kernel=Kernel.AUTO)
    rt = batch.process(ins, 512 * ch)
    assert batch.last_kernel() == Kernel.TENSOR and batch.host_pipeline_stats()[0] == 1
    for s in (0, n - 1):
        assert rt["produced"][s] == len(refs[s]["out"])
        assert np.max(np.abs(rt["out"][s].astype(np.float64) - refs[s]["out"])) <= TOL_FAST
    batch.close()


def test_config5_share_time_sliced_with_state_carry_matches_oracle():
    """BASELINE configs[4] as one GPU of eight sees it: the share of rank 3 (streams 24576..32767 of
    65 536 stereo streams x 10 s, 44.1 -> 48 kHz, 128 taps) fed as 11 device-resident time slices
    (10 x 86 calls + a tail) with the stream state carried from slice to slice -- the loop of
    bench.py's strong-scaling leg.  Counts of ALL 8192 streams and every sample of the streams with
    id = 0 mod 64 are compared with the oracle run over the unsliced 10 s."""
    from resampler_b200.fir import FLAG_ASYNC, MEM_DEVICE, DeviceBuffer, _ptr_array, _size_array
    from resampler_b200 import _lib
    from resampler_b200.sharding import shard_range
    lib = _lib.load()
    lo, hi = shard_range(65536, 8, 3)
    n, ch, in_hz, out_hz, lat, call = hi - lo, 2, 44100, 48000, 3, 512
    frames_total, slice_frames = 441000, 86 * 512
    n_full, tail = divmod(frames_total, slice_frames)
    batch = FirBatch(n, ch, in_hz, out_hz, Latency(lat), Attenuation.Db90, kernel=Kernel.AUTO)
    in_stride = slice_frames * ch
    out_stride = ((int(slice_frames / batch.ratio()) + 8) * ch + 3) & ~3
    d_in, d_out = DeviceBuffer(0, n * in_stride), DeviceBuffer(0, n * out_stride)
    assert lib.rsb_fill_synthetic(0, d_in.ptr, lo, n, slice_frames, ch, in_hz, 0x5EED) == 0
    in_ptrs = _ptr_array([d_in.ptr + 4 * s * in_stride for s in range(n)])
    out_ptrs = _ptr_array([d_out.ptr + 4 * s * out_stride for s in range(n)])
    caps = _size_array([out_stride] * n)
    check = [s for s in range(n) if (lo + s) % 64 == 0]
    got = {s: [] for s in check}
    cons_t, prod_t, calls_t = np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n, np.int64)
    for k in range(n_full + 1):
        lens = _size_array([(slice_frames if k < n_full else tail) * ch] * n)
        c, p, nc = batch.process_ptrs(in_ptrs, lens, call * ch, 0, out_ptrs, caps, memspace=MEM_DEVICE,
                                      flags=FLAG_ASYNC)
        batch.sync()
        assert batch.last_kernel() == Kernel.TENSOR
        cons_t += np.array(c[:], np.int64)
        prod_t += np.array(p[:], np.int64)
        calls_t += np.array(nc[:], np.int64)
        for s in check:
            got[s].append(d_out.download(int(p[s]), s * out_stride))
    worst = 0.0
    for s in check:
        one = d_in.download(in_stride, s * in_stride)
        x = np.concatenate([one] * n_full + [one[:tail * ch]])
        ref = oracle_stream(ch, in_hz, out_hz, lat, 1, x, call * ch)
        if s == check[0]:
            assert (cons_t == frames_total * ch).all() and (prod_t == len(ref["out"])).all()
            assert (calls_t == len(ref["consumed"])).all()
        y = np.concatenate(got[s])
        assert len(y) == len(ref["out"])
        worst = max(worst, float(np.max(np.abs(y.astype(np.float64) - ref["out"]))))
    assert worst <= TOL_FAST, worst
    batch.close()
    d_in.free()
    d_out.free()


@pytest.mark.parametrize("ch,in_hz,out_hz,lat,att,sizes", [
    (2, 44100, 48000, 3, 1, [0, 1, 100, 512, 513, 1024]),     # 128 taps stereo
    (2, 48000, 44100, 3, 1, [512]),                           # BASELINE config 1 call shape
    (8, 96000, 48000, 2, 1, [0, 64, 512, 700]),               # 64 taps, 8 channels, ratio 2
    (3, 44100, 48000, 0, 0, [7, 100, 333]),                   # odd channel count, 16 taps
    (1, 48000, 8000, 1, 2, [160, 480, 960]),                  # ratio 6
    (1, 8000, 48000, 3, 1, [1, 80, 160]),                     # ratio 1/6, 128 taps
    (2, 1000003, 999983, 2, 1, [512, 4096, 5000]),            # near-unity arbitrary rates, > capacity offers
    # integer-ratio runs (tp_run_block): phase cycle 3 / 2 / 1, sliding (D = 1) and strided windows
    (2, 16000, 48000, 1, 1, [45, 160, 333, 480]),             # cycle 3, stereo, 32 taps
    (1, 32000, 48000, 2, 1, [100, 320, 479]),                 # cycle 3, windows 2 frames apart, 64 taps
    (2, 44100, 88200, 3, 2, [256, 441]),                      # cycle 2, 128 taps
    (4, 48000, 32000, 0, 0, [96, 480]),                       # cycle 2, windows 3 apart, 4 channels, 16 taps
    (1, 48000, 48000, 1, 0, [100, 480]),                      # ratio 1: one phase
    (3, 16000, 48000, 0, 0, [50, 200]),                       # cycle 3 x 3 channels: 9 lanes per group, 5 spare lanes
    (5, 96000, 48000, 1, 1, [64, 300]),                       # 5 channels: 6 groups, 2 spare lanes
    (16, 48000, 24000, 0, 0, [128, 250]),                     # 16 channels: two groups per warp
])
def test_fused_submit_configurations_bit_exact(ch, in_hz, out_hz, lat, att, sizes):
    """The single-launch submit kernels (thread-per-output variant, fir_submit.cu) over channel
    counts, tap counts and ratios: every call's (consumed, produced) and every sample equal the
    oracle bit for bit, per-stream call sizes differing, small output capacities included."""
    from resampler_b200.fir import MEM_DEVICE, DeviceBuffer
    n = 20
    rng = np.random.default_rng(ch * 1000 + lat)
    batch = FirBatch(n, ch, in_hz, out_hz, Latency(lat), Attenuation(att), kernel=Kernel.AUTO)
    refs = [O.OracleFir(ch, in_hz, out_hz, lat, att) for _ in range(n)]
    bso = batch.buffer_size_output()
    mx = max(sizes)
    d_in = DeviceBuffer(0, n * max(mx, 1) * ch)
    d_out = DeviceBuffer(0, n * bso)
    for it in range(10):
        sz = [int(rng.choice(sizes)) for _ in range(n)]
        # every third round offers a small output buffer (capacity-limited calls, :542-545)
        cap = [bso if (it % 3 or s % 2) else 24 * ch for s in range(n)]
        ins = [noise(rng, s * ch) for s in sz]
        for s in range(n):
            if sz[s]:
                d_in.upload(ins[s], s * mx * ch)
        cons, prod = batch.submit_ptrs([d_in.ptr + 4 * s * mx * ch for s in range(n)], [x.size for x in ins],
                                       [d_out.ptr + 4 * s * bso for s in range(n)], cap, memspace=MEM_DEVICE)
        assert batch.last_kernel() == Kernel.EXACT
        for s in range(n):
            o = np.zeros(cap[s], np.float32)
            _, c, p = refs[s].resample(ins[s], o)
            assert (cons[s], prod[s]) == (c, p), (it, s)
            if p:
                assert np.array_equal(bits(d_out.download(p, s * bso)), bits(o[:p])), (it, s)
    batch.close()
    d_in.free()
    d_out.free()


def test_strong_upsampling_single_call_is_bounded_by_the_output_buffer():
    """ADVICE: the plan workspace of a single resample() call is bounded by the caller's output
    capacity, not by what 4096 buffered frames could yield: 8 Hz -> 48 kHz (ratio 1.7e-4) used to be
    refused as 'too many frames'; the reference produces min(capacity, available) frames."""
    ch, in_hz, out_hz, lat = 1, 8, 48000, 0
    r = ResamplerFir.new_from_hz(ch, in_hz, out_hz, Latency(lat), Attenuation.Db60)
    ref = O.OracleFir(ch, in_hz, out_hz, lat, 0)
    rng = np.random.default_rng(3)
    for n_in, cap in ((64, 4096), (0, 1000), (5, 12194)):
        x = noise(rng, n_in)
        o1, o2 = np.zeros(cap, np.float32), np.zeros(cap, np.float32)
        c, p = r.resample(x, o1)
        _, c2, p2 = ref.resample(x, o2)
        assert (c, p) == (c2, p2)
        assert np.array_equal(bits(o1[:p]), bits(o2[:p2]))


def test_device_side_filter_design_is_bit_identical_to_the_host_design():
    """the Kaiser-windowed-sinc table designed on the GPU (f64 Bessel window, a
    restatement of glibc's sinf, the reference's sequential f32 normalisation sum; window.rs:17-131)
    equals the host design -- and through it the oracle's table -- bit for bit, for every tap count
    and attenuation and for up-/down-sampling cutoffs; a resampler created with the device design
    switched on produces bit-identical samples (EXACT kernel)."""
    from resampler_b200.fir import device_design_table, host_design_table, set_device_filter_design
    worst_ms = 0.0
    for in_hz, out_hz in ((44100, 48000), (48000, 44100), (96000, 48000), (16000, 48000), (1000003, 999983),
                          (48000, 8000)):
        for lat in (0, 1, 2, 3):
            for att in (0, 1, 2):
                host, _ = host_design_table(in_hz, out_hz, Latency(lat), Attenuation(att))
                dev, ms = device_design_table(in_hz, out_hz, Latency(lat), Attenuation(att))
                worst_ms = max(worst_ms, ms)
                assert np.array_equal(bits(dev), bits(host)), (in_hz, out_hz, lat, att,
                                                                int(np.sum(bits(dev) != bits(host))))
    assert worst_ms < 50.0, worst_ms      # generous: the bound also has to hold under compute-sanitizer
    # a rate pair nobody has designed yet, created with the device design on
    prev = set_device_filter_design(True)
    try:
        ch, in_hz, out_hz, lat = 2, 37800, 48000, 2
        r = ResamplerFir.new_from_hz(ch, in_hz, out_hz, Latency(lat), Attenuation.Db120)
        ref = O.OracleFir(ch, in_hz, out_hz, lat, 2)
        rng = np.random.default_rng(5)
        for _ in range(4):
            x = noise(rng, 700 * ch)
            o1, o2 = np.zeros(r.buffer_size_output(), np.float32), np.zeros(r.buffer_size_output(), np.float32)
            c, p = r.resample(x, o1)
            _, c2, p2 = ref.resample(x, o2)
            assert (c, p) == (c2, p2)
            assert np.array_equal(bits(o1[:p]), bits(o2[:p2]))
        r.close()
    finally:
        set_device_filter_design(prev)


def test_pcie_probe_reports_plausible_rates():
    """rsb_pcie_probe (the ceiling bench.py holds the end-to-end figure against): every mode returns
    rates of a PCIe link, and bad arguments are refused."""
    import ctypes as C
    from resampler_b200 import _lib
    lib = _lib.load()
    out = (C.c_double * 2)()
    for mode, up, dn in ((0, True, False), (1, False, True), (2, True, True), (3, True, True)):
        assert lib.rsb_pcie_probe(0, 64 << 20, 2, mode, out) == 0
        assert (1.0 < out[0] < 200.0) == up and (1.0 < out[1] < 200.0) == dn, (mode, out[0], out[1])
    assert lib.rsb_pcie_probe(0, 0, 2, 0, out) != 0 and lib.rsb_pcie_probe(0, 1 << 20, 2, 7, out) != 0


def test_host_pipeline_async_chained_calls_equal_synchronous_ones():
    """RSB_FLAG_ASYNC on pipelined host-memspace calls (what bench.py's e2e leg does): two chained
    calls followed by sync() give the counts and samples of two synchronous calls, bit for bit
    (exact kernel), from pinned buffers."""
    import ctypes as C
    from resampler_b200 import _lib
    from resampler_b200.fir import FLAG_ASYNC, MEM_HOST
    lib = _lib.load()
    n, ch, frames = 40, 2, 220_000          # 70 MB of input: above the pipeline threshold
    bso_total = int(frames * 48000 / 44100) + 4400
    in_vals, out_vals = frames * ch, bso_total * ch
    h_in = lib.rsb_alloc_pinned(n * in_vals * 4)
    h_out = [lib.rsb_alloc_pinned(n * out_vals * 4) for _ in range(2)]
    src = np.ctypeslib.as_array(C.cast(h_in, C.POINTER(C.c_float)), shape=(n, in_vals))
    src[:] = np.random.default_rng(12).uniform(-1, 1, (n, in_vals)).astype(np.float32)
    results = {}
    for mode, flags in (("async", FLAG_ASYNC), ("sync", 0)):
        b = FirBatch(n, ch, 44100, 48000, Latency.Sample64, Attenuation.Db90, kernel=Kernel.EXACT)
        outs, counts = [], []
        for k in range(2):
            cons, prod, calls = b.process_ptrs([h_in + 4 * s * in_vals for s in range(n)], [in_vals] * n, 512 * ch, 0,
                                               [h_out[k] + 4 * s * out_vals for s in range(n)], [out_vals] * n,
                                               memspace=MEM_HOST, flags=flags)
            counts.append((list(cons[:]), list(prod[:]), list(calls[:])))
        b.sync()
        assert b.host_pipeline_stats()[0] == 2
        for k in range(2):
            a = np.ctypeslib.as_array(C.cast(h_out[k], C.POINTER(C.c_float)), shape=(n, out_vals))
            outs.append(np.array([a[s, :counts[k][1][s]].copy() for s in range(n)]))
        results[mode] = (counts, outs)
        b.close()
    assert results["async"][0] == results["sync"][0]
    for k in range(2):
        assert np.array_equal(results["async"][1][k].view(np.uint32), results["sync"][1][k].view(np.uint32))
    lib.rsb_free_pinned(h_in)
    for p in h_out:
        lib.rsb_free_pinned(p)
