"""Host-side mirror of the reference's ``ResamplerFft`` (src/resampler_fft.rs:43-246) over the C ABI
(``rsb_fft_*`` in include/resampler_b200.h): fixed-size chunks, overlap-add through a pair of FFTs on
the GPU (csrc/fft_resampler.cu).  ``FftBatch`` is the batched multi-stream form that is new."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from .fir import MEM_DEVICE, MEM_HOST, ResampleError, _ptr_array, _size_array


def _check(rc: int) -> None:
    if rc == 0:
        return
    if rc in (1, 2):
        raise ResampleError(rc)
    msg = _lib.load().rsb_fft_last_error().decode()
    if rc == 12:
        raise ValueError(msg)
    raise RuntimeError(f"resampler_b200 error {rc}: {msg}")


class FftBatch:
    """N independent ``ResamplerFft`` streams with identical parameters on one GPU."""

    def __init__(self, n_streams: int, channels: int, input_rate_hz: int, output_rate_hz: int, device: int = 0):
        self._lib = _lib.load()
        h = C.c_void_p()
        _check(self._lib.rsb_fft_create(C.byref(h), device, n_streams, channels, int(input_rate_hz),
                                        int(output_rate_hz)))
        self._h = h
        self.n_streams, self.channels = n_streams, channels

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.rsb_fft_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def chunk_size_input(self) -> int:
        return self._lib.rsb_fft_chunk_size_input(self._h)

    def chunk_size_output(self) -> int:
        return self._lib.rsb_fft_chunk_size_output(self._h)

    def delay(self) -> int:
        return self._lib.rsb_fft_delay(self._h)

    def reset(self, stream: int = -1) -> None:
        _check(self._lib.rsb_fft_reset(self._h, stream))

    def sync(self) -> None:
        _check(self._lib.rsb_fft_sync(self._h))

    def launch_count(self) -> int:
        return self._lib.rsb_fft_launch_count(self._h)

    def resample(self, stream: int, input: np.ndarray, output: np.ndarray) -> None:
        assert input.dtype == np.float32 and output.dtype == np.float32
        assert input.flags.c_contiguous and output.flags.c_contiguous
        _check(self._lib.rsb_fft_resample(self._h, stream, input.ctypes.data_as(_lib.f32p), input.size,
                                          output.ctypes.data_as(_lib.f32p), output.size))

    def process_ptrs(self, in_ptrs, in_lens, out_ptrs, out_lens, streams=None, memspace=MEM_DEVICE, flags: int = 0):
        n = len(in_ptrs)
        done = (C.c_size_t * n)()
        st = (C.c_uint32 * n)(*streams) if streams is not None else None
        _check(self._lib.rsb_fft_process_batch(self._h, n, st, _ptr_array(in_ptrs), _size_array(in_lens),
                                               _ptr_array(out_ptrs), _size_array(out_lens), done, memspace, flags))
        return done

    def process(self, inputs: Sequence[np.ndarray], streams: Optional[Sequence[int]] = None):
        """Whole chunks of every input (host arrays) -> list of output arrays."""
        csi, cso = self.chunk_size_input(), self.chunk_size_output()
        for a in inputs:
            if not isinstance(a, np.ndarray) or a.dtype != np.float32 or not a.flags["C_CONTIGUOUS"]:
                raise TypeError("process() needs C-contiguous float32 arrays")
        outs = [np.zeros((a.size // csi) * cso, np.float32) for a in inputs]
        done = self.process_ptrs([a.ctypes.data for a in inputs], [a.size for a in inputs],
                                 [o.ctypes.data if o.size else 0 for o in outs], [o.size for o in outs],
                                 streams=streams, memspace=MEM_HOST)
        return [o[:done[i] * cso] for i, o in enumerate(outs)]


class ResamplerFft:
    """Drop-in mirror of the reference's single-stream ``ResamplerFft``."""

    def __init__(self, channels: int, sample_rate_input: int, sample_rate_output: int, device: int = 0):
        self._b = FftBatch(1, channels, int(sample_rate_input), int(sample_rate_output), device)

    @classmethod
    def new(cls, channels: int, sample_rate_input: int, sample_rate_output: int):
        return cls(channels, sample_rate_input, sample_rate_output)

    def chunk_size_input(self) -> int:
        return self._b.chunk_size_input()

    def chunk_size_output(self) -> int:
        return self._b.chunk_size_output()

    def delay(self) -> int:
        return self._b.delay()

    def resample(self, input: np.ndarray, output: np.ndarray) -> None:
        """Raises ResampleError like the reference (resampler_fft.rs:186-192)."""
        self._b.resample(0, input, output)

    def close(self) -> None:
        self._b.close()
