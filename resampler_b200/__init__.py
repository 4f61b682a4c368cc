"""resampler_b200 -- B200-native (sm_100a) implementation of hasenbanck/resampler's
``ResamplerFir`` hot path.

Host-side mirror of the reference's public interface (src/lib.rs:160-163):
``ResamplerFir``, ``Latency``, ``Attenuation``, ``SampleRate``, ``ResampleError``,
plus the batched multi-stream ``FirBatch`` that is new.  All compute goes through
the C ABI of ``include/resampler_b200.h`` into hand-written CUDA kernels.
"""
from .fir import (Attenuation, FirBatch, Kernel, Latency, PcmFormat, ResampleError, ResamplerFir,
                  SampleRate, device_count)
from .fft import FftBatch, ResamplerFft

__all__ = ["Attenuation", "FftBatch", "FirBatch", "Kernel", "Latency", "PcmFormat", "ResampleError",
           "ResamplerFft", "ResamplerFir", "SampleRate", "device_count"]
__version__ = "0.1.0"
