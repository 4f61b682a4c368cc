"""GPU parity of the FFT resampler (csrc/fft_resampler.cu) against oracle/fft_oracle.py through the C
ABI.  Tolerance: the oracle transforms in f64, the kernel in f32 (as the reference does): the bar is
FFT round-off, 2e-5 absolute for |x| <= 1 (observed ~2e-6), stated here because north_star's 1e-6 is
the FIR path's."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "tests"))
import fft_oracle as F  # noqa: E402
from resampler_b200 import FftBatch, ResampleError, ResamplerFft  # noqa: E402

pytestmark = pytest.mark.gpu
TOL_FFT = 2e-5


@pytest.mark.parametrize("ch,i,o", [(1, 44100, 48000), (2, 48000, 44100), (2, 44100, 48000), (1, 16000, 48000),
                                    (2, 48000, 32000), (1, 96000, 48000), (3, 48000, 96000), (1, 16000, 44100),
                                    (2, 22050, 16000), (1, 48000, 48000), (8, 44100, 48000)])
def test_single_stream_chunks_match_the_oracle_with_state_carry(ch, i, o):
    r = ResamplerFft.new(ch, i, o)
    ref = F.OracleFft(ch, i, o)
    assert (r.chunk_size_input(), r.chunk_size_output(), r.delay()) == (ref.chunk_size_input(),
                                                                        ref.chunk_size_output(), ref.delay())
    rng = np.random.default_rng(ch * 7 + i // 1000)
    worst = 0.0
    for _ in range(5):
        x = rng.uniform(-1, 1, r.chunk_size_input()).astype(np.float32)
        y1 = np.zeros(r.chunk_size_output(), np.float32)
        y2 = np.zeros(r.chunk_size_output(), np.float32)
        r.resample(x, y1)
        assert ref.resample(x, y2) == 0
        worst = max(worst, float(np.max(np.abs(y1.astype(np.float64) - y2))))
    assert worst <= TOL_FFT, worst
    # error precedence of resampler_fft.rs:186-192
    with pytest.raises(ResampleError) as e:
        r.resample(x[:-1], y1[:-1])
    assert e.value.kind == ResampleError.InvalidInputBufferSize
    with pytest.raises(ResampleError) as e:
        r.resample(x, y1[:-1])
    assert e.value.kind == ResampleError.InvalidOutputBufferSize
    r.close()


def test_unsupported_rates_are_refused():
    with pytest.raises(ValueError):
        FftBatch(1, 1, 44100, 12345)


def test_batch_of_streams_many_chunks_equals_chunk_by_chunk_and_reference_amplitude_tests():
    """Batched entry: several chunks per stream in one launch, different lengths per stream; equals
    the oracle fed chunk by chunk.  Then the reference's own DC test (:436-470) on the GPU."""
    n, ch, i, o = 12, 2, 44100, 48000
    b = FftBatch(n, ch, i, o)
    csi, cso = b.chunk_size_input(), b.chunk_size_output()
    rng = np.random.default_rng(4)
    for rnd in range(2):                        # second round: state carried from the first
        ins = [rng.uniform(-1, 1, csi * (1 + (s + rnd) % 4) + (s % 3) * 5).astype(np.float32) for s in range(n)]
        outs = b.process(ins)
        if rnd == 0:
            refs = [F.OracleFft(ch, i, o) for _ in range(n)]
        for s in range(n):
            chunks = ins[s].size // csi
            assert outs[s].size == chunks * cso
            want = np.zeros(chunks * cso, np.float32)
            for k in range(chunks):
                refs[s].resample(ins[s][k * csi:(k + 1) * csi], want[k * cso:(k + 1) * cso])
            assert np.max(np.abs(outs[s].astype(np.float64) - want)) <= TOL_FFT, (rnd, s)
    b.reset(-1)
    x = [np.full(csi, 0.5, np.float32) for _ in range(n)]
    for _ in range(5):
        y = b.process(x)
    lo, hi = min(b.delay(), cso // 8) * 2, cso * 3 // 4
    assert np.all(np.abs(y[0][lo:hi] - 0.5) < 0.02)
    b.close()


def test_device_memspace_async_equals_host_memspace_and_reset_clears_the_overlap():
    """The batched entry on device-resident buffers (RSB_FLAG_ASYNC + sync, what bench.py times) gives
    the host-memspace result bit for bit; reset() clears the carried overlap (a second pass from a
    reset handle reproduces the first)."""
    from resampler_b200.fir import FLAG_ASYNC, MEM_DEVICE, DeviceBuffer
    n, ch, i, o, chunks = 8, 2, 48000, 44100, 5
    b = FftBatch(n, ch, i, o)
    csi, cso = b.chunk_size_input(), b.chunk_size_output()
    rng = np.random.default_rng(8)
    ins = [rng.uniform(-1, 1, csi * chunks).astype(np.float32) for _ in range(n)]
    host = b.process(ins)
    b.reset(-1)
    d_in, d_out = DeviceBuffer(0, n * csi * chunks), DeviceBuffer(0, n * cso * chunks)
    for s in range(n):
        d_in.upload(ins[s], s * csi * chunks)
    for _ in range(2):                                  # second round after reset(): same result
        done = b.process_ptrs([d_in.ptr + 4 * s * csi * chunks for s in range(n)], [csi * chunks] * n,
                              [d_out.ptr + 4 * s * cso * chunks for s in range(n)], [cso * chunks] * n,
                              memspace=MEM_DEVICE, flags=FLAG_ASYNC)
        b.sync()
        assert list(done[:]) == [chunks] * n
        for s in range(n):
            got = d_out.download(cso * chunks, s * cso * chunks)
            assert np.array_equal(got.view(np.uint32), host[s].view(np.uint32)), s
        b.reset(-1)
    b.close()
    d_in.free()
    d_out.free()
