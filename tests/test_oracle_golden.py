"""Pins the CPU oracle (oracle/fir_oracle.c) before anything trusts it.

Every known-answer vector the reference's own tests hold for this path is
replayed here (file:line cited per test), followed by the characterisation
vectors of SURVEY.md section 8(a) and a bit-for-bit cross-check of the C
restatement against the independent Python restatement oracle/plan_numpy.py.
CPU only.
"""
import importlib.util
import random
import sys
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as O

ROOT = Path(__file__).resolve().parent.parent
_spec = importlib.util.spec_from_file_location("plan_numpy", ROOT / "oracle" / "plan_numpy.py")
plan_numpy = importlib.util.module_from_spec(_spec)
sys.modules["plan_numpy"] = plan_numpy
_spec.loader.exec_module(plan_numpy)


def approx_rel(actual, expected, tol):
    return abs(actual / expected - 1.0) < tol


# ---- src/window.rs:152-160 ----
@pytest.mark.parametrize("x,expected", [
    (0.0, 1.000000000000000), (1.0, 1.266065877752008), (2.0, 2.279585302336067),
    (5.0, 27.239871823604442), (10.0, 2815.716628466254)])
def test_bessel_i0_known_values(x, expected):
    assert approx_rel(O.bessel_i0(x), expected, 1e-6)


# ---- src/window.rs:162-228 (periodic) and :296-362 (symmetric) ----
KAISER_CASES = [
    (5, 0.5, False, [0.940306193319157, 0.978296239370539, 0.997576503537205,
                     0.997576503537205, 0.978296239370539]),
    (15, 5.0, False, [0.036710892271287, 0.120260370289032, 0.248940523358684,
                      0.414903639243367, 0.599303856150336, 0.775322104445407,
                      0.913812483869200, 0.990113103661532, 0.990113103661532,
                      0.913812483869200, 0.775322104445407, 0.599303856150336,
                      0.414903639243367, 0.248940523358684, 0.120260370289032]),
    (9, 10.0, False, [0.000355149374724, 0.030999213508099, 0.203914483842615,
                      0.581810162428082, 0.942963979134466, 0.942963979134466,
                      0.581810162428082, 0.203914483842615, 0.030999213508099]),
    (5, 0.5, True, [0.940306193319157, 0.984902269883833, 1.0, 0.984902269883833,
                    0.940306193319157]),
    (15, 5.0, True, [0.036710892271287, 0.127982199301765, 0.270694417889417,
                     0.453689854203301, 0.651738235245363, 0.830535847455841,
                     0.955247316456437, 1.0, 0.955247316456437, 0.830535847455841,
                     0.651738235245363, 0.453689854203301, 0.270694417889417,
                     0.127982199301765, 0.036710892271287]),
    (9, 10.0, True, [0.000355149374724, 0.041939800327748, 0.282059620822733,
                     0.740117133443384, 1.0, 0.740117133443384, 0.282059620822733,
                     0.041939800327748, 0.000355149374724]),
]


@pytest.mark.parametrize("n,beta,sym,expected", KAISER_CASES)
def test_kaiser_window_scipy_values(n, beta, sym, expected):
    w = O.kaiser_window(n, beta, sym)
    assert len(w) == len(expected)
    for a, e in zip(w, expected):
        assert approx_rel(float(a), e, 1e-5)


# ---- src/window.rs:230-237 ----
@pytest.mark.parametrize("n,expected", [
    (64, 0.8999482371370552), (128, 0.9499741185685276), (256, 0.9749870592842638),
    (512, 0.9874935296421319), (1024, 0.9937467648210659)])
def test_cutoff_kaiser_values(n, expected):
    assert approx_rel(O.cutoff_kaiser(n, 10.0), expected, 1e-6)


# ---- src/window.rs:239-247 ----
def test_cutoff_kaiser_valid_range():
    for n in (32, 64, 128, 256, 512, 1024, 2048):
        c = O.cutoff_kaiser(n, 10.0)
        assert 0.0 < c < 1.0


# ---- src/window.rs:249-271 ----
def test_make_sincs_dimensions():
    t, _ = O.make_sincs(4, 2, 0.9, 10.0, False)
    assert t.shape == (2, 4)


# ---- src/window.rs:273-294 and :364-385 ----
@pytest.mark.parametrize("sym,expected", [
    (False, [[-0.0084796025, 0.4976338439, 0.4976338439, -0.0084796025],
             [-0.0000355271, 0.0296676259, 0.9623917926, 0.0296676259]]),
    (True, [[-0.0135119673, 0.6818196469, 0.3016755841, -0.0000802533],
            [-0.0000397065, 0.0471924586, 0.9759149497, 0.0070292878]])])
def test_make_sincs_reference_values(sym, expected):
    t, _ = O.make_sincs(4, 2, 0.9, 10.0, sym)
    for row, erow in zip(t, expected):
        for a, e in zip(row, erow):
            assert approx_rel(float(a), e, 1e-5)


# ---- src/window.rs:387-410 ----
def test_make_sincs_normalization():
    t, _ = O.make_sincs(8, 4, 0.95, 10.0, False)
    total = np.float32(0)
    for row in t:
        total = np.float32(total + np.float32(row.sum(dtype=np.float32)))
    assert abs(float(total) - 4.0) < 0.01


# ---- src/fir/mod.rs:137-192: SIMD order vs scalar within 1e-5 ----
@pytest.mark.parametrize("kind", [O.CONV_AVX512, O.CONV_AVX512_INTRIN])
def test_convolve_simd_vs_scalar(kind):
    if kind == O.CONV_AVX512_INTRIN and not O.lib().orc_cpu_has_avx512f():
        pytest.skip("host has no avx512f")
    for taps in (16, 32, 64, 128):
        for frac in (0.0, 0.25, 0.5, 0.75, 1.0):
            i = np.arange(taps, dtype=np.float32)
            x = np.sin(i * np.float32(0.1)).astype(np.float32)
            c1 = np.cos(i * np.float32(0.2)).astype(np.float32)
            c2 = np.cos(i * np.float32(0.15)).astype(np.float32)
            s = O.convolve(x, c1, c2, frac, O.CONV_SCALAR)
            v = O.convolve(x, c1, c2, frac, kind)
            assert abs(float(s) - float(v)) < 1e-5, (taps, frac)
    taps = 64
    imp = np.zeros(taps, np.float32)
    imp[0] = 1.0
    c1 = np.arange(taps, dtype=np.float32)
    c2 = (np.arange(taps) * 2).astype(np.float32)
    s = O.convolve(imp, c1, c2, 0.3, O.CONV_SCALAR)
    v = O.convolve(imp, c1, c2, 0.3, kind)
    assert abs(float(s) - float(v)) < 1e-5
    assert float(v) == 0.0


def test_avx512_intrinsics_match_lanewise_restatement_bitwise():
    """The scalar-coded 16-lane restatement IS the AVX-512 routine: bit-identical results."""
    if not O.lib().orc_cpu_has_avx512f():
        pytest.skip("host has no avx512f")
    rng = np.random.default_rng(7)
    tab = O.design_table(44100, 48000, 3, 1)
    for _ in range(2000):
        x = rng.uniform(-1, 1, 128).astype(np.float32)
        p = int(rng.integers(0, 1023))
        f = np.float32(rng.uniform(0, 1))
        a = O.convolve(x, tab[p], tab[p + 1], f, O.CONV_AVX512)
        b = O.convolve(x, tab[p], tab[p + 1], f, O.CONV_AVX512_INTRIN)
        assert a.view(np.uint32) == b.view(np.uint32)


# ---- src/resampler_fir.rs:741-815: stop-band attenuation >= 90 dB ----
def _impulse_response(in_hz, out_hz, duration=5.0):
    n = int(np.float32(in_hz) * np.float32(duration))
    pos = int(min(np.float32(n) * np.float32(0.5), np.float32(n) - np.float32(1.0)))
    x = np.zeros(n, np.float32)
    x[pos] = 1.0
    f = O.OracleFir(1, in_hz, out_hz, O.LATENCY["Sample64"], O.ATTENUATION["Db90"])
    return f.process(x, 256)["out"]


@pytest.mark.parametrize("in_hz,out_hz", [(22050, 44100), (22050, 48000)])
def test_stopband_attenuation(in_hz, out_hz):
    y = _impulse_response(in_hz, out_hz)
    peak = int(np.argmax(np.abs(y)))
    win = int(np.float32(out_hz) * np.float32(0.1))
    start = max(peak - win // 2, 0)
    end = min(start + win, len(y))
    ir = y[start:end]
    fft_size = 8192
    buf = np.zeros(fft_size, np.float32)
    m = min(len(ir), fft_size)
    buf[:m] = ir[:m]
    mag = np.abs(np.fft.rfft(buf.astype(np.float64)))
    db = np.where(mag > 1e-10, 20 * np.log10(np.maximum(mag, 1e-300)), -200.0)

    def to_bin(fhz):
        return int(round(fhz / out_hz * fft_size))

    nyq = in_hz / 2.0
    pb = db[to_bin(20.0):to_bin(nyq * 0.9) + 1]
    sb_end = min(len(db) - 10, to_bin(out_hz / 2.0 * 0.95))
    sb = db[to_bin(nyq * 1.1):sb_end + 1]
    att = pb.max() - sb.max()
    assert att >= 90.0, att
    # SURVEY.md 8(a) [probe]: 108.61 dB / 109.42 dB with an independent restatement
    expected = {44100: 108.61, 48000: 109.42}[out_hz]
    assert abs(att - expected) < 0.05, att


# ---- src/resampler_fir.rs:817-839, :841-850, :852-862 ----
def test_new_from_hz_matches_new_and_arbitrary_rates_and_zero_rates():
    a = O.OracleFir(1, 48000, 44100, 3, 1)
    b = O.OracleFir(1, 48000, 44100, 3, 1)
    x = np.full(512, 0.5, np.float32)
    oa = np.zeros(a.buffer_size_output(), np.float32)
    ob = np.zeros(b.buffer_size_output(), np.float32)
    ra = a.resample(x, oa)
    rb = b.resample(x, ob)
    assert ra == rb and np.array_equal(oa, ob)
    c = O.OracleFir(1, 24000, 16000, O.LATENCY["Sample32"], O.ATTENUATION["Db60"])
    err, _, _ = c.resample(np.zeros(256, np.float32), np.zeros(c.buffer_size_output(), np.float32))
    assert err == 0
    with pytest.raises(ValueError):
        O.OracleFir(1, 0, 44100)
    with pytest.raises(ValueError):
        O.OracleFir(1, 44100, 0)


def test_error_precedence():
    """resampler_fir.rs:514-519: input length is checked before output length."""
    f = O.OracleFir(2, 48000, 44100, 3, 1)
    assert f.resample(np.zeros(3, np.float32), np.zeros(3, np.float32))[0] == 1
    assert f.resample(np.zeros(4, np.float32), np.zeros(3, np.float32))[0] == 2
    assert f.resample(np.zeros(4, np.float32), np.zeros(4, np.float32))[0] == 0


# ---- SURVEY.md 8(c) [probe] constants ----
def test_design_probe_constants():
    cases = [((44100, 48000, 128), 0x3F733181, 1.0526561737060547),
             ((48000, 44100, 128), 0x3F5F6F15, 1.1456893682479858),
             ((16000, 48000, 32), 0x3F4CC604, 1.2501579523086548),
             ((96000, 48000, 64), 0x3EE66302, 2.222421646118164)]
    for (i, o, t), bits, s in cases:
        c = O.cutoff_for(i, o, t, 10.0)
        assert int(c.view(np.uint32)) == bits
        _, ssum = O.make_sincs(t, 1024, c, 10.0, True)
        # glibc sinf dependent: equal on glibc 2.39; tolerate a last-ulp libm difference
        assert abs(float(ssum) - s) <= 2.4e-7 * s
    tab = O.design_table(44100, 48000, 3, 1)
    assert np.allclose(tab[0, :3], [-1.5854984667384997e-06, 1.8371410988038406e-06,
                                    -1.462913246541575e-06], rtol=1e-5, atol=0)
    assert np.allclose(tab[512, 63:65], [0.6333794593811035, 0.6355674266815186], rtol=1e-6)
    for ph, rs in ((0, 1.0000033), (512, 1.0000052), (1023, 1.0000031)):
        assert abs(tab[ph].sum(dtype=np.float64) - rs) < 5e-7


def test_buffer_size_output_values():
    """SURVEY.md 8(a) row 9 (resampler_fir.rs:456-465)."""
    cases = [((2, 48000, 44100, 3), 7296), ((2, 44100, 48000, 3), 8642),
             ((1, 16000, 48000, 1), 12194), ((8, 96000, 48000, 2), 16144)]
    for (ch, i, o, lat), exp in cases:
        assert O.OracleFir(ch, i, o, lat, 1).buffer_size_output() == exp


# ---- SURVEY.md 8(a) characterisation vectors (counts) ----
COUNT_CASES = [
    # in_hz, out_hz, latency, call_frames, first produced frames, steady leftover
    (44100, 48000, 3, 512, [420, 557, 557, 557, 558, 557], 127),
    (48000, 44100, 3, 512, [354, 471, 470, 470, 471], 127),
    (16000, 48000, 1, 160, [388, 480, 480, 480, 479, 480], 31),
    (16000, 48000, 1, 64, [100, 191, 192, 192, 192], 31),
    (96000, 48000, 2, 512, [225, 256, 256], 62),
]


@pytest.mark.parametrize("in_hz,out_hz,lat,call,first,leftover", COUNT_CASES)
def test_count_vectors(in_hz, out_hz, lat, call, first, leftover):
    f = O.OracleFir(1, in_hz, out_hz, lat, 1)
    x = np.zeros(call * 40, np.float32)
    r = f.process(x, call)
    assert list(r["produced"][:len(first)]) == first
    assert np.all(r["consumed"] == call)
    p = plan_numpy.PlanFir(in_hz, out_hz, lat)
    for _ in range(40):
        p.call(call, 10 ** 9)
    assert p.available == leftover


def test_count_totals_60s():
    """60 s totals: calls and output frames."""
    for (i, o, lat, call, n_calls, n_out) in [
            (44100, 48000, 3, 512, 5168, 2879862), (48000, 44100, 3, 512, 5625, 2645884),
            (16000, 48000, 1, 160, 6000, 2879907), (96000, 48000, 2, 512, 11250, 2879969)]:
        f = O.OracleFir(1, i, o, lat, 1)
        r = f.process(np.zeros(i * 60, np.float32), call)
        assert r["calls"] == n_calls and len(r["out"]) == n_out


def test_max_call_and_small_capacity_vectors():
    f = O.OracleFir(1, 44100, 48000, 3, 1)
    x = np.zeros(44100 * 4, np.float32)
    r = f.process(x, 4096)
    assert list(r["consumed"][:3]) == [4096, 3969, 3969]
    assert list(r["produced"][:3]) == [4321, 4320, 4320]
    # output capacity of 100 frames: exercises :553 and :528
    f = O.OracleFir(1, 44100, 48000, 3, 1)
    r = f.process(x, 512, out_cap_len=100)
    assert list(r["consumed"][:12]) == [512] * 9 + [314, 92, 92]
    # 384 kHz -> 1 kHz, 16 taps: the 0.5.1 lookahead fix (:593-596)
    f = O.OracleFir(1, 384000, 1000, 0, 1)
    r = f.process(np.zeros(512 * 12, np.float32), 512)
    assert list(r["produced"][:6]) == [2, 1, 1, 2, 1, 1]


def test_plan_vectors_first_call():
    """SURVEY.md 8(a) per-output plan vectors, 44.1->48 T128, first 512-frame call."""
    f = O.OracleFir(1, 44100, 48000, 3, 1)
    out = np.zeros(f.buffer_size_output(), np.float32)
    err, c, p, tr = f.resample(np.zeros(512, np.float32), out, trace=True)
    assert (err, c, p) == (0, 512, 420)
    expect = {0: (0, 0, 1, 0x00000000), 1: (0, 940, 941, 0x3F4CCCCD),
              2: (1, 857, 858, 0x3F19999A), 3: (2, 774, 775, 0x3ECCCCCD),
              4: (3, 691, 692, 0x3E4CCCCD), 5: (4, 608, 609, 0), 10: (9, 192, 193, 0),
              15: (13, 799, 800, 0x3F800000), 160: (147, 0, 1, 0x2E800000),
              161: (147, 940, 941, 0x3F4CCCCD), 418: (384, 38, 39, 0x3ECCCCCD),
              419: (384, 979, 980, 0x3E4CCCCD)}
    for k, (off, p1, p2, fb) in expect.items():
        assert (int(tr["input_offset"][k]), int(tr["phase1"][k]), int(tr["phase2"][k]),
                int(tr["frac_bits"][k])) == (off, p1, p2, fb), k


# ---- C restatement vs independent Python restatement, bit for bit ----
RATE_PAIRS = [(44100, 48000), (48000, 44100), (16000, 48000), (96000, 48000), (192000, 8000),
              (8000, 192000), (384000, 1000), (44100, 44101), (22050, 48000), (48000, 48000)]


@pytest.mark.parametrize("in_hz,out_hz", RATE_PAIRS)
def test_c_oracle_matches_python_restatement(in_hz, out_hz):
    rnd = random.Random(in_hz * 31 + out_hz)
    for lat in (0, 3):
        f = O.OracleFir(1, in_hz, out_hz, lat, 1)
        p = plan_numpy.PlanFir(in_hz, out_hz, lat)
        assert f.buffer_size_output() == p.buffer_size_output_frames()
        for _ in range(120):
            n_in = rnd.choice([0, 1, 7, 64, 160, 512, 1000, 4096, 9000, rnd.randrange(0, 9000)])
            cap = rnd.choice([0, 1, 100, 5000, 20000, rnd.randrange(0, 20000)])
            out = np.zeros(max(cap, 1), np.float32)[:cap]
            err, c, pr, tr = f.resample(np.zeros(n_in, np.float32), out, trace=True)
            to_copy, outs = p.call(n_in, cap)
            assert err == 0 and c == to_copy and pr == len(outs)
            if outs:
                arr = np.array(outs, dtype=np.uint64)
                assert np.array_equal(arr[:, 0], tr["input_offset"])
                assert np.array_equal(arr[:, 1], tr["phase1"])
                assert np.array_equal(arr[:, 2], tr["phase2"])
                assert np.array_equal(arr[:, 3], tr["frac_bits"])


# ---------------------------------------------------------------------------------------------
# CLI format step (resample/src/main.rs:128-156) -- the reference has no test for it; the
# expected values are the exact quotients `s as f32 / 2^(bits-1)` worked out by hand
# ---------------------------------------------------------------------------------------------
def test_pcm_format_step_known_answers():
    s16 = np.array([-32768, 32767, 0, 1, -1, 16384], np.int16)
    assert np.array_equal(O.pcm_to_f32(s16, O.PCM_S16),
                          np.array([-1.0, 32767 / 32768, 0.0, 2.0 ** -15, -(2.0 ** -15), 0.5],
                                   np.float32))
    u8 = np.array([0, 128, 255, 192], np.uint8)        # hound: unsigned on disk, minus 128
    assert np.array_equal(O.pcm_to_f32(u8, O.PCM_U8),
                          np.array([-1.0, 0.0, 127 / 128, 0.5], np.float32))
    s24 = np.array([0x00, 0x00, 0x80,  0xFF, 0xFF, 0x7F,  0x00, 0x00, 0x40,  0xFF, 0xFF, 0xFF],
                   np.uint8)                            # -2^23, 2^23-1, 2^22, -1
    assert np.array_equal(O.pcm_to_f32(s24, O.PCM_S24),
                          np.array([-1.0, (2 ** 23 - 1) / 2 ** 23, 0.5, -(2.0 ** -23)], np.float32))
    # 32 bits: `(1 << 31) as f32` is an i32 shift = -2^31, so the CLI inverts the polarity;
    # `s as f32` rounds 16777217 to 16777216 (nearest even) first
    s32 = np.array([-(1 << 31), (1 << 31) - 1, 1 << 30, 16777217], np.int32)
    assert np.array_equal(O.pcm_to_f32(s32, O.PCM_S32),
                          np.array([1.0, -1.0, -0.5, -(16777216 / 2 ** 31)], np.float32))
    f32 = np.array([0.25, -1.5, 3e-39], np.float32)     # passes through, denormal included
    assert np.array_equal(O.pcm_to_f32(f32, O.PCM_F32).view(np.uint32), f32.view(np.uint32))


def test_pcm_mono_to_stereo_duplication():
    s16 = np.array([100, -200, 300], np.int16)          # main.rs:139-146
    want = np.repeat(s16.astype(np.float32) / np.float32(32768.0), 2)
    assert np.array_equal(O.pcm_to_f32(s16, O.PCM_S16, dup=2), want)
    assert O.pcm_to_f32(np.zeros(0, np.int16), O.PCM_S16, dup=2).size == 0


# ---------------------------------------------------------------------------------------------
# committed regression fixtures (tests/golden/fir_cases.npz, made by tests/golden/make_fixtures.py)
# ---------------------------------------------------------------------------------------------
def _golden():
    import importlib.util
    from pathlib import Path
    gdir = Path(__file__).resolve().parent / "golden"
    spec = importlib.util.spec_from_file_location("make_fixtures", gdir / "make_fixtures.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, np.load(gdir / "fir_cases.npz")


def test_oracle_reproduces_committed_fixtures():
    """The oracle built here gives the frozen counts, plan entries and sample bits."""
    mod, z = _golden()
    for name in mod.CASES:
        got = mod.run_case(name)
        for key, val in got.items():
            want = z[f"{name}/{key}"]
            assert val.shape == want.shape, (name, key)
            if val.dtype == np.float32:
                assert np.array_equal(val.view(np.uint32), want.view(np.uint32)), (name, key)
            else:
                assert np.array_equal(val, want), (name, key)


def test_pcm_format_step_matches_independent_numpy_restatement():
    """A second restatement of main.rs:131-136 in numpy (`s as f32 / max_value` with the actual
    f32 division) against the C oracle on random samples of every integer format."""
    rng = np.random.default_rng(17)
    n = 20000
    s16 = rng.integers(-32768, 32768, n, dtype=np.int16)
    assert np.array_equal(O.pcm_to_f32(s16, O.PCM_S16), s16.astype(np.float32) / np.float32(32768.0))
    u8 = rng.integers(0, 256, n, dtype=np.uint8)
    assert np.array_equal(O.pcm_to_f32(u8, O.PCM_U8),
                          (u8.astype(np.int32) - 128).astype(np.float32) / np.float32(128.0))
    v24 = rng.integers(-(1 << 23), 1 << 23, n, dtype=np.int64)
    u = (v24 & 0xFFFFFF).astype(np.uint32)
    b24 = np.stack([u & 0xFF, (u >> 8) & 0xFF, (u >> 16) & 0xFF], axis=1).astype(np.uint8).reshape(-1)
    assert np.array_equal(O.pcm_to_f32(b24, O.PCM_S24), v24.astype(np.float32) / np.float32(8388608.0))
    s32 = rng.integers(-(1 << 31), 1 << 31, n, dtype=np.int64).astype(np.int32)
    # `(1 << 31) as f32` on an i32 literal = -2147483648.0
    assert np.array_equal(O.pcm_to_f32(s32, O.PCM_S32), s32.astype(np.float32) / np.float32(-2147483648.0))
