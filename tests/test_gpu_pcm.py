"""GPU parity of the CLI batch path: the format step
(resample/src/main.rs:128-156: integer PCM -> f32, mono -> stereo) fused in front of the
canonical 512-value-call loop (:226-254), through the C ABI, against the CPU oracle.

Bar: converted samples, counts and (EXACT kernel) output samples bit-identical; the
tensor / FFMA2 kernels within 1e-6 absolute.
"""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
from resampler_b200 import Attenuation, FirBatch, Kernel, Latency, PcmFormat
from resampler_b200 import _lib
from resampler_b200.fir import MEM_DEVICE

pytestmark = pytest.mark.gpu

TOL = 1e-6   # absolute, north_star


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def raw_samples(rng, fmt, n):
    """n random samples of the format, extremes included, as the bytes a WAV data chunk holds."""
    if fmt == PcmFormat.U8:
        a = rng.integers(0, 256, n, dtype=np.uint8)
        a[:2] = (0, 255)[:n]
        return a
    if fmt == PcmFormat.S16:
        a = rng.integers(-32768, 32768, n, dtype=np.int16)
        a[:2] = (-32768, 32767)[:n]
        return a
    if fmt == PcmFormat.S24:
        v = rng.integers(-(1 << 23), 1 << 23, n, dtype=np.int64)
        v[:2] = (-(1 << 23), (1 << 23) - 1)[:n]
        u = (v & 0xFFFFFF).astype(np.uint32)
        return np.stack([u & 0xFF, (u >> 8) & 0xFF, (u >> 16) & 0xFF], axis=1).astype(np.uint8).reshape(-1)
    if fmt == PcmFormat.S32:
        a = rng.integers(-(1 << 31), 1 << 31, n, dtype=np.int64).astype(np.int32)
        a[:3] = (-(1 << 31), (1 << 31) - 1, 16777217)[:n]   # the last one rounds in `as f32`
        return a
    return rng.uniform(-1.0, 1.0, n).astype(np.float32)


def oracle_cli(ch, in_hz, out_hz, lat, raw, fmt, src_ch, call_len=512):
    x = O.pcm_to_f32(raw, int(fmt), 1 if src_ch == ch else ch)
    f = O.OracleFir(ch, in_hz, out_hz, lat, 1)
    return x, f.process(x, call_len)


@pytest.mark.parametrize("fmt", list(PcmFormat))
@pytest.mark.parametrize("src_ch", [1, 2])
def test_cli_batch_path_bit_exact(fmt, src_ch):
    """Mono and stereo files of every sample format, ragged lengths (value counts that are not
    multiples of 4, one empty file), 44.1 -> 48 kHz as the CLI would run them."""
    ch = 2
    rng = np.random.default_rng(100 + int(fmt) * 2 + src_ch)
    frames = [0, 1, 3, 1021, 4099, 9001]
    raws = [raw_samples(rng, fmt, f * src_ch) for f in frames]
    batch = FirBatch(len(frames), ch, 44100, 48000, Latency.Sample64, Attenuation.Db90,
                     kernel=Kernel.EXACT)
    res = batch.process_pcm(raws, fmt, src_ch)     # call_len = 512 values, main.rs:227
    assert batch.launch_count() > 0
    for s, raw in enumerate(raws):
        _, ref = oracle_cli(ch, 44100, 48000, 3, raw, fmt, src_ch)
        assert res["consumed"][s] == ref["consumed_total"], s
        assert res["calls"][s] == ref["calls"], s
        assert res["produced"][s] == len(ref["out"]), s
        assert np.array_equal(bits(res["out"][s]), bits(ref["out"])), s
    batch.close()


def test_format_step_known_values_through_the_gpu():
    """A 1:1 rate pair with the impulse-like phase-0 row is not an identity filter, so the
    converted samples are observed directly: resample F32 and integer renderings of the same
    signal and require bit-identical outputs."""
    ch = 2
    rng = np.random.default_rng(5)
    s16 = raw_samples(rng, PcmFormat.S16, 6000 * ch)
    as_f32 = (s16.astype(np.float32) / np.float32(32768.0)).astype(np.float32)
    batch = FirBatch(2, ch, 48000, 44100, Latency.Sample64, Attenuation.Db90, kernel=Kernel.EXACT)
    a = batch.process_pcm([s16], PcmFormat.S16, ch, streams=[0])
    b = batch.process_pcm([as_f32], PcmFormat.F32, ch, streams=[1])
    assert a["produced"][0] == b["produced"][0] > 0
    assert np.array_equal(bits(a["out"][0]), bits(b["out"][0]))
    # the 32-bit polarity quirk of the reference (`(1 << 31) as f32` = -2^31) is kept
    s32 = (s16.astype(np.int32) << 16)
    batch.reset()
    c = batch.process_pcm([s32], PcmFormat.S32, ch, streams=[0])
    neg = batch.process_pcm([-as_f32], PcmFormat.F32, ch, streams=[1])
    assert np.array_equal(bits(c["out"][0]), bits(neg["out"][0]))
    batch.close()


def test_mono_source_into_eight_channels_and_state_carry():
    """Generic duplication factor (mono -> 8 channels) and a second batch that continues from
    the carried history / phase."""
    ch = 8
    rng = np.random.default_rng(9)
    raws1 = [raw_samples(rng, PcmFormat.S16, n) for n in (3001, 2999, 17)]
    raws2 = [raw_samples(rng, PcmFormat.S16, n) for n in (555, 1, 640)]
    batch = FirBatch(3, ch, 96000, 48000, Latency.Sample32, Attenuation.Db90, kernel=Kernel.EXACT)
    r1 = batch.process_pcm(raws1, PcmFormat.S16, 1, call_len=512 * ch)
    r2 = batch.process_pcm(raws2, PcmFormat.S16, 1, call_len=512 * ch)
    for s in range(3):
        f = O.OracleFir(ch, 96000, 48000, 2, 1)
        o1 = f.process(O.pcm_to_f32(raws1[s], O.PCM_S16, ch), 512 * ch)
        o2 = f.process(O.pcm_to_f32(raws2[s], O.PCM_S16, ch), 512 * ch)
        assert np.array_equal(bits(r1["out"][s]), bits(o1["out"])), s
        assert np.array_equal(bits(r2["out"][s]), bits(o2["out"])), s
    batch.close()


def test_errors():
    batch = FirBatch(2, 2, 44100, 48000, Latency.Sample64, Attenuation.Db90)
    x = np.zeros(1000, np.int16)
    with pytest.raises(ValueError, match="Unsupported channel count"):   # main.rs:151-154
        batch.process_pcm([np.zeros(999, np.int16)], PcmFormat.S16, 3)
    with pytest.raises(ValueError, match="unknown PCM format"):
        n = 1
        rc = batch._lib.rsb_fir_process_pcm_batch(
            batch._h, n, None, (C.c_void_p * n)(x.ctypes.data), (C.c_size_t * n)(500), 9, 2, 512, 0,
            (C.c_void_p * n)(x.ctypes.data), (C.c_size_t * n)(0), None, None, None, 1, 0)
        from resampler_b200.fir import _check
        _check(rc)
    from resampler_b200 import ResampleError
    with pytest.raises(ResampleError):           # call_len not a multiple of channels
        batch.process_pcm([x], PcmFormat.S16, 2, call_len=511)
    batch.close()


@pytest.mark.parametrize("kernel,fmt,ch,src_ch,aligned", [
    (Kernel.TENSOR, PcmFormat.S16, 2, 2, True),     # format step fused into the tensor kernel
    (Kernel.TENSOR, PcmFormat.S16, 2, 2, False),    # rows not 16-byte aligned: separate format pass
    (Kernel.TENSOR, PcmFormat.S16, 2, 1, True),     # fused, mono source duplicated (main.rs:139-146)
    (Kernel.TENSOR, PcmFormat.S16, 1, 1, True),     # fused, mono streams
    (Kernel.TENSOR, PcmFormat.S24, 2, 1, True),     # fused, packed 24-bit mono source
    (Kernel.TENSOR, PcmFormat.S24, 2, 2, True),     # fused, packed 24-bit stereo
    (Kernel.TENSOR, PcmFormat.S24, 1, 1, True),     # fused, packed 24-bit mono streams
    (Kernel.TENSOR, PcmFormat.S32, 2, 2, True),     # 32-bit: separate format pass
    (Kernel.FAST, PcmFormat.S16, 2, 1, True),
])
def test_device_resident_pcm_batch_fast_kernels(kernel, fmt, ch, src_ch, aligned):
    """64 equally long files resident on the device (raw bytes in HBM): the converted staging
    buffer is equally strided, so the tensor kernel takes it; s16 rows that are 16-byte aligned
    are converted inside the tensor kernel's loader.  Samples within 1e-6 of the oracle, counts
    exact; a second batch checks the history written from the raw tail."""
    n, frames = 64, 6007
    lib = _lib.load()
    rng = np.random.default_rng(77 + int(fmt))
    raws = [raw_samples(rng, fmt, frames * src_ch) for _ in range(n)]
    raws2 = [raw_samples(rng, fmt, 1500 * src_ch) for _ in range(n)]
    bps = fmt.bytes_per_sample()
    raw_bytes = frames * src_ch * bps
    stride = (raw_bytes + 15) & ~15 if aligned else raw_bytes + 6
    d_raw = lib.rsb_alloc_device(0, stride * n + 16)
    batch = FirBatch(n, ch, 44100, 48000, Latency.Sample64, Attenuation.Db90, kernel=kernel)
    cap = int(frames * ch / batch.ratio()) + 4 * batch.buffer_size_output()
    cap -= cap % 4
    d_out = lib.rsb_alloc_device(0, cap * 4 * n)
    assert d_raw and d_out

    def run(rs, nf):
        for i, r in enumerate(rs):
            b = np.ascontiguousarray(r).view(np.uint8)
            assert lib.rsb_memcpy(0, d_raw + i * stride, b.ctypes.data, b.nbytes, 0) == 0
        return batch.process_pcm_ptrs(
            [d_raw + i * stride for i in range(n)], [nf] * n, fmt, src_ch, 512, 0,
            [d_out + i * cap * 4 for i in range(n)], [cap] * n, memspace=MEM_DEVICE)

    def fetch(i, count):
        got = np.empty(count, np.float32)
        assert lib.rsb_memcpy(0, got.ctypes.data, d_out + i * cap * 4, got.nbytes, 1) == 0
        return got

    cons, prod, calls = run(raws, frames)
    assert batch.last_kernel() == kernel
    assert batch.last_ingest_ms() > 0.0
    assert batch.last_pcm_fused() == (kernel == Kernel.TENSOR and aligned
                                      and fmt in (PcmFormat.S16, PcmFormat.S24))
    firsts = {i: fetch(i, prod[i]) for i in (0, 1, 31, 63)}
    cons2, prod2, _ = run(raws2, 1500)
    worst = 0.0
    for i in (0, 1, 31, 63):
        x = O.pcm_to_f32(raws[i], int(fmt), 1 if src_ch == ch else ch)
        x2 = O.pcm_to_f32(raws2[i], int(fmt), 1 if src_ch == ch else ch)
        f = O.OracleFir(ch, 44100, 48000, 3, 1)
        ref = f.process(x, 512)
        ref2 = f.process(x2, 512)
        assert cons[i] == ref["consumed_total"] and prod[i] == len(ref["out"])
        assert calls[i] == ref["calls"]
        assert cons2[i] == ref2["consumed_total"] and prod2[i] == len(ref2["out"])
        worst = max(worst, float(np.max(np.abs(firsts[i].astype(np.float64) - ref["out"]))))
        worst = max(worst, float(np.max(np.abs(fetch(i, prod2[i]).astype(np.float64) - ref2["out"]))))
    assert worst <= TOL, worst
    lib.rsb_free_device(0, d_raw)
    lib.rsb_free_device(0, d_out)
    batch.close()


@pytest.mark.parametrize("fmt", [PcmFormat.S16, PcmFormat.S24])
@pytest.mark.parametrize("ch,src_ch", [(2, 2), (2, 1), (1, 1)])
def test_fused_and_separate_format_step_agree_bit_for_bit(monkeypatch, ch, src_ch, fmt):
    """The tensor kernel converting raw s16 frames in its loader and the separate format pass
    feed the same f32 values into the same arithmetic: identical output bits (host buffers)."""
    n, frames = 64, 5000
    rng = np.random.default_rng(123)
    raws = [raw_samples(rng, fmt, frames * src_ch) for _ in range(n)]
    outs = []
    for unfused in (False, True):
        if unfused:
            monkeypatch.setenv("RSB_PCM_UNFUSED", "1")
        batch = FirBatch(n, ch, 44100, 48000, Latency.Sample64, Attenuation.Db90, kernel=Kernel.TENSOR)
        res = batch.process_pcm(raws, fmt, src_ch, call_len=512 * ch)
        assert batch.last_pcm_fused() == (not unfused)
        tail = batch.flush()
        outs.append([np.concatenate([a, b]) for a, b in zip(res["out"], tail)])
        batch.close()
    for a, b in zip(*outs):
        assert a.size > frames * ch and np.array_equal(bits(a), bits(b))


def test_wav_files_end_to_end(tmp_path):
    """tools/resample_wav.py: a mono 16-bit and a stereo 24-bit file, different lengths, one
    call -- every output file equals what the oracle produces for that file alone (the tool
    runs the default AUTO kernel selection: two streams => the bit-exact kernel)."""
    import sys
    sys.path.insert(0, str(O.ROOT / "tools"))
    import resample_wav
    from resampler_b200.wav import read_wav, write_wav_pcm
    rng = np.random.default_rng(11)
    a = raw_samples(rng, PcmFormat.S16, 5000)
    b = raw_samples(rng, PcmFormat.S24, 2 * 3777)
    (tmp_path / "in").mkdir()
    write_wav_pcm(tmp_path / "in" / "a.wav", a, 44100, 1, 16)
    write_wav_pcm(tmp_path / "in" / "b.wav", b, 44100, 2, 24)
    done = resample_wav.resample_files([tmp_path / "in" / "a.wav", tmp_path / "in" / "b.wav"],
                                       tmp_path / "out", 48000)
    assert len(done) == 2
    for name, raw, fmt, src_ch in (("a.wav", a, PcmFormat.S16, 1), ("b.wav", b, PcmFormat.S24, 2)):
        w = read_wav(tmp_path / "out" / name)
        assert (w.sample_rate, w.channels, w.fmt) == (48000, 2, PcmFormat.F32)
        _, ref = oracle_cli(2, 44100, 48000, 3, raw, fmt, src_ch)
        assert np.array_equal(w.raw.view(np.uint32), bits(ref["out"]))


def test_larger_device_batch_fused_equals_separate_and_exact(monkeypatch):
    """512 stereo s16 files x 2 s on the device, asynchronous submits: the fused loader and the
    separate format pass agree bit for bit on EVERY sample of every stream, and both stay within
    1e-6 of the bit-exact kernel fed by the separate pass."""
    from resampler_b200.fir import FLAG_ASYNC
    n, ch, frames = 512, 2, 88200
    lib = _lib.load()
    rng = np.random.default_rng(2024)
    raw = rng.integers(-32768, 32768, (n, frames * ch), dtype=np.int16)
    stride = frames * ch * 2
    assert stride % 16 == 0
    d_raw = lib.rsb_alloc_device(0, stride * n)
    assert lib.rsb_memcpy(0, d_raw, raw.ctypes.data, raw.nbytes, 0) == 0
    cap = ((int(frames * 48000 / 44100) + 8) * ch + 3) & ~3
    d_out = lib.rsb_alloc_device(0, cap * 4 * n)
    outs = {}
    for name, kern, unfused in (("fused", Kernel.TENSOR, False), ("separate", Kernel.TENSOR, True),
                                ("exact", Kernel.EXACT, True)):
        if unfused:
            monkeypatch.setenv("RSB_PCM_UNFUSED", "1")
        b = FirBatch(n, ch, 44100, 48000, Latency.Sample64, Attenuation.Db90, kernel=kern)
        cons, prod, calls = b.process_pcm_ptrs(
            [d_raw + s * stride for s in range(n)], [frames] * n, PcmFormat.S16, ch, 512, 0,
            [d_out + s * cap * 4 for s in range(n)], [cap] * n, memspace=MEM_DEVICE, flags=FLAG_ASYNC)
        b.sync()
        assert b.last_pcm_fused() == (name == "fused")
        assert all(c == frames * ch for c in cons[:]) and len(set(prod[:])) == 1
        got = np.empty(n * cap, np.float32)
        assert lib.rsb_memcpy(0, got.ctypes.data, d_out, got.nbytes, 1) == 0
        outs[name] = got.reshape(n, cap)[:, :prod[0]].copy()
        b.close()
    assert np.array_equal(bits(outs["fused"]), bits(outs["separate"]))
    assert np.max(np.abs(outs["fused"].astype(np.float64) - outs["exact"])) <= TOL
    ref = oracle_cli(ch, 44100, 48000, 3, raw[n - 1], PcmFormat.S16, ch)[1]
    assert np.array_equal(bits(outs["exact"][n - 1]), bits(ref["out"]))
    lib.rsb_free_device(0, d_raw)
    lib.rsb_free_device(0, d_out)


@pytest.mark.parametrize("fmt", [PcmFormat.S16, PcmFormat.S24])
@pytest.mark.parametrize("ch,src_ch", [(2, 2), (2, 1), (1, 1)])
def test_fused_loader_tiny_and_odd_lengths(ch, src_ch, fmt):
    """Fused raw-s16 loader at the small end: files of 1 .. 300 frames (shorter than one TMA
    box, than the filter, not multiples of anything), several batches in a row on the same
    streams so that history and phase carry through the raw tail conversion."""
    n = 64
    rng = np.random.default_rng(900 + ch * 10 + src_ch)
    batch = FirBatch(n, ch, 44100, 48000, Latency.Sample64, Attenuation.Db90, kernel=Kernel.TENSOR)
    refs = [O.OracleFir(ch, 44100, 48000, 3, 1) for _ in range(4)]
    worst = 0.0
    for frames in (1, 5, 17, 127, 128, 300, 2):
        raws = [raw_samples(rng, fmt, frames * src_ch) for _ in range(n)]
        res = batch.process_pcm(raws, fmt, src_ch, call_len=512 * ch)
        assert batch.last_pcm_fused() and batch.last_kernel() == Kernel.TENSOR
        for k, s in enumerate((0, 1, 31, 63)):
            x = O.pcm_to_f32(raws[s], int(fmt), 1 if src_ch == ch else ch)
            ref = refs[k].process(x, 512 * ch)
            assert res["consumed"][s] == ref["consumed_total"] and res["produced"][s] == len(ref["out"])
            if len(ref["out"]):
                worst = max(worst, float(np.max(np.abs(res["out"][s].astype(np.float64) - ref["out"]))))
    assert worst <= TOL, worst
    batch.close()
