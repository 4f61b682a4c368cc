"""oracle/fft_oracle.py -- CPU restatement of the reference's `ResamplerFft` (TEST INFRASTRUCTURE ONLY:
imported by tests/ and nowhere else).

What is restated line by line: the conversion table and its scaling (src/fft/planner.rs:33-233), the
constructor's filter (src/resampler_fft.rs:349-386, through the C oracle's `orc_make_sincs` /
`orc_cutoff_kaiser`, which are pinned by the reference's own window.rs tests), the per-chunk
overlap-add loop (`FftResampler::resample`, :388-424) and the chunk API (:182-246).

What is NOT restated: the reference's own mixed-radix f32 FFT (src/fft/radix_fft.rs, real_complex/*,
4400 lines).  The transforms here are numpy's, in f64: the mathematical DFT the reference's FFT
approximates in f32.  So the oracle pins the ALGORITHM (sizes, filter, spectrum truncation, scaling,
overlap-add, state carry) and is exact where the reference is within f32 FFT round-off of it;
sample parity against it is therefore stated with an FFT round-off tolerance (tests/test_gpu_fft.py),
**parity unpinned by reference sample values** -- the reference's tests for this path hold none, only
amplitude checks with EPSILON = 0.02 (:427-566), which tests/test_fft_oracle.py replays.
"""
from __future__ import annotations

import numpy as np

KAISER_BETA = 10.0   # src/resampler_fft.rs:16

# src/lib.rs:191-204 (family), :219-233 (Hz)
FAMILY = {22050: 22050, 16000: 16000, 32000: 16000, 44100: 22050, 48000: 48000, 88200: 22050,
          96000: 48000, 176400: 22050, 192000: 48000, 384000: 48000}

# src/fft/planner.rs:46-155: (input family, output family) -> (base in, base out)
BASE = {(48000, 48000): (2, 2), (22050, 22050): (2, 2), (16000, 16000): (2, 2),
        (22050, 48000): (588, 1280), (48000, 22050): (1280, 588),
        (16000, 48000): (64, 192), (48000, 16000): (192, 64),
        (16000, 22050): (640, 882), (22050, 16000): (882, 640)}


def conversion_sizes(in_hz: int, out_hz: int):
    """(fft_size_input, fft_size_output) as `ResamplerFft::new` uses them (resampler_fft.rs:82-84):
    base sizes x family multipliers (planner.rs:158-160), then scaled by a power of two to at least
    512 input samples (planner.rs:211-224)."""
    if in_hz not in FAMILY or out_hz not in FAMILY:
        raise ValueError("ResamplerFft takes the SampleRate enum's rates only")
    fi, fo = FAMILY[in_hz], FAMILY[out_hz]
    base_in, base_out = BASE[(fi, fo)]
    size_in, size_out = base_in * (in_hz // fi), base_out * (out_hz // fo)
    mult = max(1, int(np.ceil(np.float32(512.0) / np.float32(size_in))))
    mult = 1 << (mult - 1).bit_length()                      # next_power_of_two
    return size_in * mult, size_out * mult


class OracleFft:
    def __init__(self, channels: int, in_hz: int, out_hz: int):
        import oracle_lib as O
        self.channels = channels
        self.n_in, self.n_out = conversion_sizes(in_hz, out_hz)
        n_in, n_out = self.n_in, self.n_out
        # resampler_fft.rs:362-368
        if n_in > n_out:
            cutoff = O.cutoff_kaiser(n_out, KAISER_BETA) * (n_out / n_in)
        else:
            cutoff = O.cutoff_kaiser(n_in, KAISER_BETA)
        sincs, _ = O.make_sincs(n_in, 1, np.float32(cutoff), KAISER_BETA, False)    # Periodic window, :370-376
        filter_time = np.zeros(2 * n_in, np.float32)
        filter_time[:n_in] = sincs[0] / np.float32(2 * n_in)                        # :380-382 (f32 division)
        self.filter_time = filter_time
        self.filter_spectrum = np.fft.rfft(filter_time.astype(np.float64))          # :384-385, n_in + 1 bins
        self.overlaps = np.zeros((channels, n_out), np.float64)                     # :86
        self.new_length = n_in + 1 if n_in < n_out else n_out                       # :399-402

    def chunk_size_input(self) -> int:
        return self.n_in * self.channels                                            # :88

    def chunk_size_output(self) -> int:
        return self.n_out * self.channels                                           # :89

    def delay(self) -> int:
        return self.n_in // 2                                                       # :151-153

    def reset(self) -> None:
        self.overlaps[:] = 0.0

    def resample(self, inp: np.ndarray, out: np.ndarray) -> int:
        """One chunk (:182-246).  Returns 0, or the reference's error (1 input / 2 output size)."""
        if len(inp) < self.chunk_size_input():
            return 1
        if len(out) < self.chunk_size_output():
            return 2
        n_in, n_out, ch = self.n_in, self.n_out, self.channels
        x = np.asarray(inp[:n_in * ch], np.float32).reshape(n_in, ch)
        for c in range(ch):
            buf = np.zeros(2 * n_in, np.float64)
            buf[:n_in] = x[:, c]                                                    # :390-391
            spec = np.fft.rfft(buf)                                                 # :393-397
            o_spec = np.zeros(n_out + 1, np.complex128)
            nl = self.new_length
            o_spec[:nl] = spec[:nl] * self.filter_spectrum[:nl]                     # :404-411
            y = np.fft.irfft(o_spec, 2 * n_out) * (2 * n_out)                       # :413-417, unnormalised
            out[c:n_out * ch:ch] = (y[:n_out] + self.overlaps[c]).astype(np.float32)   # :419-422
            self.overlaps[c] = y[n_out:]                                            # :423
        return 0
