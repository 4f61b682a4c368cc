"""The FFT-resampler oracle (oracle/fft_oracle.py) against everything the reference's own tests hold
for this path: the conversion table (src/fft/planner.rs:236-300), and the amplitude-preservation
tests of src/resampler_fft.rs:427-566 (EPSILON = 0.02)."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
import fft_oracle as F  # noqa: E402

EPSILON = 0.02


def test_conversion_table_known_answers():
    # planner.rs tests: base sizes before scaling = scaled sizes / power of two
    cases = {(48000, 96000): (2, 4), (48000, 192000): (2, 8), (22050, 48000): (588, 1280),
             (16000, 48000): (64, 192), (16000, 44100): (640, 1764), (44100, 48000): (1176, 1280)}
    for (i, o), (bi, bo) in cases.items():
        n_in, n_out = F.conversion_sizes(i, o)
        mult = n_in // bi
        assert n_in == bi * mult and n_out == bo * mult and mult & (mult - 1) == 0
        assert n_in >= 512 and (mult == 1 or bi * mult // 2 < 512)
    assert F.conversion_sizes(44100, 48000) == (1176, 1280)
    assert F.conversion_sizes(48000, 44100) == (1280, 1176)
    assert F.conversion_sizes(48000, 96000) == (512, 1024)
    assert F.conversion_sizes(16000, 48000) == (512, 1536)
    assert F.conversion_sizes(48000, 32000) == (768, 512)
    with pytest.raises(ValueError):
        F.conversion_sizes(44100, 12345)


@pytest.mark.parametrize("i,o", [(48000, 44100), (44100, 48000), (48000, 32000), (32000, 48000), (96000, 48000),
                                 (48000, 96000)])
def test_dc_signal_amplitude_preservation(i, o):
    r = F.OracleFft(1, i, o)
    x = np.full(r.chunk_size_input(), 0.5, np.float32)
    y = np.zeros(r.chunk_size_output(), np.float32)
    for _ in range(5):
        assert r.resample(x, y) == 0
    a, b = min(r.delay(), len(y) // 4), len(y) * 3 // 4
    assert np.all(np.abs(y[a:b] - 0.5) < EPSILON)


@pytest.mark.parametrize("i,o", [(48000, 44100), (44100, 48000), (48000, 32000)])
def test_sine_wave_amplitude_preservation(i, o):
    r = F.OracleFft(1, i, o)
    n = r.chunk_size_input()
    phase = np.float32(0.0)
    inc = np.float32(2.0 * np.pi * 1000.0 / i)
    x = np.empty(n, np.float32)
    for k in range(n):                       # the reference accumulates the phase in f32
        x[k] = np.float32(0.5) * np.sin(phase)
        phase = np.float32(phase + inc)
    y = np.zeros(r.chunk_size_output(), np.float32)
    for _ in range(5):
        r.resample(x, y)
    a, b = min(r.delay(), len(y) // 4), len(y) * 3 // 4
    assert abs(float(np.max(np.abs(y[a:b]))) - 0.5) < EPSILON


def test_stereo_dc_and_errors():
    r = F.OracleFft(2, 48000, 44100)
    n = r.chunk_size_input()
    x = np.zeros(n, np.float32)
    x[0::2], x[1::2] = 0.3, 0.6
    y = np.zeros(r.chunk_size_output(), np.float32)
    for _ in range(5):
        r.resample(x, y)
    a, b = min(r.delay(), len(y) // 8) * 2, len(y) * 3 // 4
    assert np.all(np.abs(y[a:b:2] - 0.3) < EPSILON) and np.all(np.abs(y[a + 1:b:2] - 0.6) < EPSILON)
    assert r.resample(x[:-1], y) == 1 and r.resample(x, y[:-1]) == 2          # :186-192
    assert (r.chunk_size_input(), r.chunk_size_output(), r.delay()) == (2560, 2352, 640)
