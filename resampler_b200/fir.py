"""Host-side mirror of the reference's ``ResamplerFir`` interface over the C ABI.

Names, argument meaning and error behaviour follow hasenbanck/resampler v0.5.1:
``ResamplerFir::new / new_from_hz / buffer_size_output / resample / delay / reset``
(src/resampler_fir.rs:252, 295, 456, 509, 630, 638), ``Latency`` (:138-162),
``Attenuation`` (:101-123), ``SampleRate`` (src/lib.rs:166-254) and
``ResampleError`` (src/error.rs:1-26).  ``FirBatch`` is the new batched entry.
"""
from __future__ import annotations

import ctypes as C
import enum
from typing import Optional, Sequence

import numpy as np

from . import _lib


class SampleRate(enum.IntEnum):
    """src/lib.rs:166-188; the value is the rate in Hz (``u32::from``, lib.rs:219-234)."""
    Hz16000 = 16000
    Hz22050 = 22050
    Hz32000 = 32000
    Hz44100 = 44100
    Hz48000 = 48000
    Hz88200 = 88200
    Hz96000 = 96000
    Hz176400 = 176400
    Hz192000 = 192000
    Hz384000 = 384000

    @classmethod
    def try_from(cls, hz: int) -> "SampleRate":
        """``TryFrom<u32>`` (lib.rs:236-254): raises ValueError for any other rate."""
        return cls(hz)


class Latency(enum.IntEnum):
    """src/resampler_fir.rs:138-162; default Sample64 (:147)."""
    Sample8 = 0
    Sample16 = 1
    Sample32 = 2
    Sample64 = 3

    def taps(self) -> int:
        return (16, 32, 64, 128)[int(self)]

    @classmethod
    def default(cls) -> "Latency":
        return cls.Sample64


class Attenuation(enum.IntEnum):
    """src/resampler_fir.rs:101-123; default Db120 (:108)."""
    Db60 = 0
    Db90 = 1
    Db120 = 2

    @classmethod
    def default(cls) -> "Attenuation":
        return cls.Db120


class Kernel(enum.IntEnum):
    AUTO = 0
    EXACT = 1   # bit-identical to the reference's AVX-512 summation order
    FAST = 2    # pre-interpolated rows, within 1e-6 absolute
    TENSOR = 3  # tcgen05 tensor cores, 3xTF32, within 1e-6 absolute (uniform batches)


class PcmFormat(enum.IntEnum):
    """Raw sample formats of the CLI's format step (resample/src/main.rs:128-137)."""
    U8 = 0
    S16 = 1
    S24 = 2    # 3 packed little-endian bytes per sample
    S32 = 3    # the reference's `(1 << 31) as f32` quirk: comes out with inverted polarity
    F32 = 4

    def bytes_per_sample(self) -> int:
        return (1, 2, 3, 4, 4)[int(self)]


_PCM_DTYPES = {PcmFormat.U8: np.uint8, PcmFormat.S16: np.int16, PcmFormat.S24: np.uint8,
               PcmFormat.S32: np.int32, PcmFormat.F32: np.float32}


class ResampleError(Exception):
    """src/error.rs:1-26.  ``kind`` is ``InvalidInputBufferSize`` or ``InvalidOutputBufferSize``."""

    InvalidInputBufferSize = 1
    InvalidOutputBufferSize = 2

    def __init__(self, kind: int):
        self.kind = kind
        msg = {1: "Input buffer size is invalid", 2: "Output buffer size is invalid"}[kind]
        super().__init__(msg)


MEM_DEVICE, MEM_HOST = 0, 1
FLAG_ASYNC, FLAG_RECORD_CALLS, FLAG_KEEP_PLAN = 1, 2, 4


def device_count() -> int:
    return _lib.load().rsb_device_count()


def _check(rc: int) -> None:
    if rc == 0:
        return
    if rc in (1, 2):
        raise ResampleError(rc)
    lib = _lib.load()
    msg = lib.rsb_last_error().decode() or lib.rsb_status_string(rc).decode()
    if rc in (10, 11):
        # the reference panics with exactly these texts (resampler_fir.rs:302-309)
        raise ValueError(lib.rsb_status_string(rc).decode())
    if rc == 12:
        raise ValueError(msg)
    rec = (C.c_uint32 * 32)()
    lib.rsb_debug_tc_hang(rec)
    if rec[0]:     # the tensor kernel's barrier watchdog fired: say where
        warps = {w: (rec[4 + w] & 0xff, (rec[4 + w] >> 8) & 0xf, rec[4 + w] >> 12) for w in range(24)
                 if rec[4 + w]}
        msg += (f" [tensor-kernel watchdog: wait tag {rec[0]}, block {rec[1]}, warp {rec[2]}, "
                f"parity {rec[3] & 0xff}, barrier smem address {rec[3] >> 8:#x}; stuck warps "
                f"(tag, parity, barrier): {warps}]")
    raise RuntimeError(f"resampler_b200 error {rc}: {msg}")


def _ptr_array(ptrs: Sequence[int]):
    if isinstance(ptrs, C.Array):      # already a ctypes array (hot loops build it once)
        return ptrs
    return (C.c_void_p * len(ptrs))(*ptrs)


def _size_array(vals: Sequence[int]):
    if isinstance(vals, C.Array):
        return vals
    return (C.c_size_t * len(vals))(*[int(v) for v in vals])


class DeviceBuffer:
    """A plain f32 device allocation made through the C ABI (no torch needed)."""

    def __init__(self, device: int, n_values: int):
        self.device = device
        self.n_values = int(n_values)
        self.ptr = _lib.load().rsb_alloc_device(device, max(self.n_values, 1) * 4)
        if not self.ptr:
            raise MemoryError(_lib.last_error())

    def upload(self, a: np.ndarray, offset_values: int = 0) -> None:
        a = np.ascontiguousarray(a, np.float32)
        assert offset_values + a.size <= self.n_values
        _check(_lib.load().rsb_memcpy(self.device, self.ptr + offset_values * 4,
                                      a.ctypes.data, a.nbytes, 0))

    def download(self, n_values: Optional[int] = None, offset_values: int = 0) -> np.ndarray:
        n = self.n_values - offset_values if n_values is None else int(n_values)
        out = np.empty(n, np.float32)
        if n:
            _check(_lib.load().rsb_memcpy(self.device, out.ctypes.data,
                                          self.ptr + offset_values * 4, n * 4, 1))
        return out

    def free(self) -> None:
        if getattr(self, "ptr", None):
            _lib.load().rsb_free_device(self.device, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class FirBatch:
    """N independent ``ResamplerFir`` streams with identical parameters on one GPU."""

    def __init__(self, n_streams: int, channels: int, input_rate_hz: int, output_rate_hz: int,
                 latency: Latency = Latency.Sample64, attenuation: Attenuation = Attenuation.Db120,
                 device: int = 0, kernel: Kernel = Kernel.AUTO):
        self._lib = _lib.load()
        h = C.c_void_p()
        _check(self._lib.rsb_fir_create(C.byref(h), device, n_streams, channels,
                                        int(input_rate_hz), int(output_rate_hz), int(latency),
                                        int(attenuation)))
        self._h = h
        self.n_streams = n_streams
        self.channels = channels
        self.device = device
        if kernel != Kernel.AUTO:
            self.set_kernel(kernel)

    # -- lifecycle ---------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.rsb_fir_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- reference interface -------------------------------------------------
    def set_kernel(self, kernel: Kernel) -> None:
        _check(self._lib.rsb_fir_set_kernel(self._h, int(kernel)))

    def last_kernel(self) -> Kernel:
        """Kernel the most recent batch actually ran on (``Kernel.AUTO`` before the first)."""
        return Kernel(self._lib.rsb_fir_last_kernel(self._h))

    def buffer_size_output(self) -> int:
        return self._lib.rsb_fir_buffer_size_output(self._h)

    def delay(self) -> int:
        return self._lib.rsb_fir_delay(self._h)

    def taps(self) -> int:
        return self._lib.rsb_fir_taps(self._h)

    def ratio(self) -> float:
        return self._lib.rsb_fir_ratio(self._h)

    def reset(self, stream: int = -1) -> None:
        _check(self._lib.rsb_fir_reset(self._h, stream))

    def coeffs(self) -> np.ndarray:
        out = np.empty((1024, self.taps()), np.float32)
        _check(self._lib.rsb_fir_coeffs(self._h, out.ctypes.data_as(_lib.f32p), out.size))
        return out

    def resample(self, stream: int, input: np.ndarray, output: np.ndarray):
        """One ``resample()`` call on one stream with host slices -> (consumed, produced)."""
        assert input.dtype == np.float32 and output.dtype == np.float32
        assert input.flags.c_contiguous and output.flags.c_contiguous
        c, p = C.c_size_t(0), C.c_size_t(0)
        _check(self._lib.rsb_fir_resample(self._h, stream, input.ctypes.data_as(_lib.f32p),
                                          input.size, output.ctypes.data_as(_lib.f32p),
                                          output.size, C.byref(c), C.byref(p)))
        return c.value, p.value

    # -- batched -------------------------------------------------------------
    def submit(self, inputs: Sequence[np.ndarray], outputs: Sequence[np.ndarray],
               streams: Optional[Sequence[int]] = None, flags: int = 0):
        """n independent ``resample()`` calls (host arrays).  Returns (consumed[], produced[])."""
        n = len(inputs)
        assert len(outputs) == n
        for a in list(inputs) + list(outputs):
            assert a.dtype == np.float32 and a.flags.c_contiguous
        cons, prod = (C.c_size_t * n)(), (C.c_size_t * n)()
        st = (C.c_uint32 * n)(*streams) if streams is not None else None
        _check(self._lib.rsb_fir_submit_batch(
            self._h, n, st, _ptr_array([a.ctypes.data for a in inputs]),
            _size_array([a.size for a in inputs]),
            _ptr_array([a.ctypes.data for a in outputs]), _size_array([a.size for a in outputs]),
            cons, prod, MEM_HOST, flags))
        return np.array(cons[:], np.int64), np.array(prod[:], np.int64)

    def submit_ptrs(self, in_ptrs, in_lens, out_ptrs, out_lens, streams=None, memspace=MEM_DEVICE,
                    flags: int = 0):
        n = len(in_ptrs)
        cons, prod = (C.c_size_t * n)(), (C.c_size_t * n)()
        st = (C.c_uint32 * n)(*streams) if streams is not None else None
        _check(self._lib.rsb_fir_submit_batch(self._h, n, st, _ptr_array(in_ptrs),
                                              _size_array(in_lens), _ptr_array(out_ptrs),
                                              _size_array(out_lens), cons, prod, memspace, flags))
        self._hold((cons, prod))
        return cons, prod

    def process(self, inputs: Sequence[np.ndarray], call_len: int, out_cap_len: int = 0,
                out_capacity: Optional[int] = None, streams: Optional[Sequence[int]] = None,
                flags: int = 0):
        """Canonical caller loop per stream (host arrays).  Returns a dict with ``out`` (list of
        arrays), ``consumed``, ``produced`` (values) and ``calls`` per stream."""
        n = len(inputs)
        for a in inputs:     # the C ABI reads raw f32 memory: no silent reinterpretation
            if not isinstance(a, np.ndarray) or a.dtype != np.float32 or not a.flags["C_CONTIGUOUS"]:
                raise TypeError("process() needs C-contiguous float32 arrays")
        if out_capacity is None:
            longest = max((a.size for a in inputs), default=0)
            out_capacity = int(longest / self.ratio()) + 4 * self.buffer_size_output() + 64
            out_capacity -= out_capacity % self.channels
        outs = [np.zeros(out_capacity, np.float32) for _ in range(n)]
        cons, prod, calls = (C.c_size_t * n)(), (C.c_size_t * n)(), (C.c_uint32 * n)()
        st = (C.c_uint32 * n)(*streams) if streams is not None else None
        _check(self._lib.rsb_fir_process_batch(
            self._h, n, st, _ptr_array([a.ctypes.data for a in inputs]),
            _size_array([a.size for a in inputs]), call_len, out_cap_len,
            _ptr_array([a.ctypes.data for a in outs]), _size_array([out_capacity] * n),
            cons, prod, calls, MEM_HOST, flags))
        return {"out": [o[:prod[i]] for i, o in enumerate(outs)],
                "consumed": np.array(cons[:], np.int64), "produced": np.array(prod[:], np.int64),
                "calls": np.array(calls[:], np.int64)}

    def process_pcm(self, inputs: Sequence[np.ndarray], fmt: PcmFormat, src_channels: int,
                    call_len: int = 512, out_cap_len: int = 0,
                    out_capacity: Optional[int] = None, streams: Optional[Sequence[int]] = None,
                    flags: int = 0):
        """The CLI's batch path (resample/src/main.rs:128-156 + 226-254) for many files at once:
        ``inputs[i]`` holds one file's raw interleaved samples (dtype of ``fmt``; S24 as bytes),
        converted to f32 and, for a mono source, duplicated into every channel ON THE GPU, then
        run through the canonical 512-value-call loop.  Same result dict as ``process``."""
        fmt = PcmFormat(fmt)
        n = len(inputs)
        bps = fmt.bytes_per_sample()
        # uint8 arrays are taken as the file's raw bytes, anything else must already have the
        # format's sample type (a silent value cast would change the audio)
        inputs = [np.asarray(a) for a in inputs]
        for a in inputs:
            if a.dtype != np.uint8 and a.dtype != _PCM_DTYPES[fmt]:
                raise TypeError(f"{fmt.name} input must be uint8 bytes or {np.dtype(_PCM_DTYPES[fmt]).name}, "
                                f"not {a.dtype.name}")
        inputs = [np.ascontiguousarray(a) for a in inputs]
        frames = []
        for a in inputs:
            if a.nbytes % (bps * src_channels) != 0:
                raise ResampleError(1)      # not a whole number of source frames
            frames.append(a.nbytes // (bps * src_channels))
        if out_capacity is None:
            longest = max(frames, default=0) * self.channels
            out_capacity = int(longest / self.ratio()) + 4 * self.buffer_size_output() + 64
            out_capacity -= out_capacity % self.channels
        outs = [np.zeros(out_capacity, np.float32) for _ in range(n)]
        cons, prod, calls = (C.c_size_t * n)(), (C.c_size_t * n)(), (C.c_uint32 * n)()
        st = (C.c_uint32 * n)(*streams) if streams is not None else None
        _check(self._lib.rsb_fir_process_pcm_batch(
            self._h, n, st, _ptr_array([a.ctypes.data for a in inputs]), _size_array(frames),
            int(fmt), src_channels, call_len, out_cap_len,
            _ptr_array([a.ctypes.data for a in outs]), _size_array([out_capacity] * n),
            cons, prod, calls, MEM_HOST, flags))
        return {"out": [o[:prod[i]] for i, o in enumerate(outs)],
                "consumed": np.array(cons[:], np.int64), "produced": np.array(prod[:], np.int64),
                "calls": np.array(calls[:], np.int64)}

    def process_pcm_ptrs(self, in_ptrs, in_frames, fmt: PcmFormat, src_channels: int, call_len,
                         out_cap_len, out_ptrs, out_capacities, streams=None,
                         memspace=MEM_DEVICE, flags: int = 0):
        n = len(in_ptrs)
        cons, prod, calls = (C.c_size_t * n)(), (C.c_size_t * n)(), (C.c_uint32 * n)()
        st = (C.c_uint32 * n)(*streams) if streams is not None else None
        _check(self._lib.rsb_fir_process_pcm_batch(
            self._h, n, st, _ptr_array(in_ptrs), _size_array(in_frames), int(fmt), src_channels,
            call_len, out_cap_len, _ptr_array(out_ptrs), _size_array(out_capacities), cons, prod,
            calls, memspace, flags))
        self._hold((cons, prod, calls))
        return cons, prod, calls

    def flush(self, streams: Optional[Sequence[int]] = None):
        """Opt-in tail handling (not in the reference): feeds ``delay()`` frames of silence to
        the listed streams (default: all) and returns what that produces, one array each."""
        n = len(streams) if streams is not None else self.n_streams
        cap = self.buffer_size_output()
        outs = [np.zeros(cap, np.float32) for _ in range(n)]
        prod = (C.c_size_t * n)()
        st = (C.c_uint32 * n)(*streams) if streams is not None else None
        _check(self._lib.rsb_fir_flush_batch(self._h, n, st, _ptr_array([a.ctypes.data for a in outs]),
                                             _size_array([cap] * n), prod, MEM_HOST, 0))
        return [o[:prod[i]] for i, o in enumerate(outs)]

    def last_pcm_fused(self) -> bool:
        """True when the last PCM batch was converted inside the tensor kernel's loader."""
        return bool(self._lib.rsb_fir_last_pcm_fused(self._h))

    def last_ingest_ms(self) -> float:
        """Device time of the format-step kernel of the most recent PCM batch."""
        ms = C.c_float(0)
        _check(self._lib.rsb_fir_last_ingest_ms(self._h, C.byref(ms)))
        return ms.value

    def process_ptrs(self, in_ptrs, total_lens, call_len, out_cap_len, out_ptrs, out_capacities,
                     streams=None, memspace=MEM_DEVICE, flags: int = 0, want_counts: bool = True):
        n = len(in_ptrs)
        cons = (C.c_size_t * n)() if want_counts else None
        prod = (C.c_size_t * n)() if want_counts else None
        calls = (C.c_uint32 * n)() if want_counts else None
        st = (C.c_uint32 * n)(*streams) if streams is not None else None
        _check(self._lib.rsb_fir_process_batch(self._h, n, st, _ptr_array(in_ptrs),
                                               _size_array(total_lens), call_len, out_cap_len,
                                               _ptr_array(out_ptrs), _size_array(out_capacities),
                                               cons, prod, calls, memspace, flags))
        self._hold((cons, prod, calls))
        return cons, prod, calls

    def _hold(self, arrays) -> None:
        """Count arrays of an async call are written when a later call (or sync) completes it:
        keep the most recent ones alive (four submits can be in flight)."""
        keep = getattr(self, "_keep", None)
        if keep is None:
            keep = self._keep = []
        keep.append(arrays)
        del keep[:-8]

    def sync(self) -> None:
        _check(self._lib.rsb_fir_sync(self._h))

    def last_call_counts(self, job: int, max_calls: int = 1 << 20):
        n = C.c_size_t(0)
        _check(self._lib.rsb_fir_last_call_counts(self._h, job, None, None, 0, C.byref(n)))
        k = min(n.value, max_calls)
        c = np.zeros(max(k, 1), np.uint32)
        p = np.zeros(max(k, 1), np.uint32)
        _check(self._lib.rsb_fir_last_call_counts(self._h, job, c.ctypes.data_as(_lib.u32p),
                                                  p.ctypes.data_as(_lib.u32p), k, C.byref(n)))
        return c[:k], p[:k]

    def last_plan(self, job: int):
        n = C.c_size_t(0)
        _check(self._lib.rsb_fir_last_plan(self._h, job, None, None, None, None, 0, C.byref(n)))
        k = n.value
        arrs = [np.zeros(max(k, 1), np.uint32) for _ in range(4)]
        _check(self._lib.rsb_fir_last_plan(self._h, job, *[a.ctypes.data_as(_lib.u32p) for a in arrs],
                                           k, C.byref(n)))
        return {"input_offset": arrs[0][:k], "phase1": arrs[1][:k], "phase2": arrs[2][:k],
                "frac_bits": arrs[3][:k]}

    # -- timing --------------------------------------------------------------
    def timer_start(self) -> None:
        _check(self._lib.rsb_fir_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float(0)
        _check(self._lib.rsb_fir_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def conv_times_ms(self, max_n: int = 64) -> np.ndarray:
        """Device time of the convolution kernel of the most recent batches, newest first."""
        ms = np.zeros(max_n, np.float32)
        n = C.c_size_t(0)
        _check(self._lib.rsb_fir_conv_times(self._h, ms.ctypes.data_as(_lib.f32p), max_n,
                                            C.byref(n)))
        return ms[:n.value]

    def launch_count(self) -> int:
        return self._lib.rsb_fir_launch_count(self._h)

    def host_pipeline_stats(self):
        """(batches, slices) of host-memspace ``process`` calls that ran as an overlapped pipeline
        of time slices (H2D / kernels / D2H), see ``rsb_fir_host_pipeline_stats``."""
        b, s = C.c_uint64(0), C.c_uint64(0)
        _check(self._lib.rsb_fir_host_pipeline_stats(self._h, C.byref(b), C.byref(s)))
        return int(b.value), int(s.value)


class ResamplerFir:
    """Drop-in mirror of the reference's single-stream ``ResamplerFir``.

    ``ResamplerFir(channels, input_rate, output_rate, latency, attenuation)`` corresponds to
    ``ResamplerFir::new`` (src/resampler_fir.rs:252-266); ``new_from_hz`` to :295-404.
    """

    def __init__(self, channels: int, input_rate: SampleRate, output_rate: SampleRate,
                 latency: Latency = Latency.Sample64, attenuation: Attenuation = Attenuation.Db120,
                 device: int = 0, kernel: Kernel = Kernel.AUTO):
        self._b = FirBatch(1, channels, int(input_rate), int(output_rate), latency, attenuation,
                           device, kernel)
        self.channels = channels

    @classmethod
    def new(cls, channels, input_rate: SampleRate, output_rate: SampleRate,
            latency: Latency = Latency.Sample64, attenuation: Attenuation = Attenuation.Db120,
            **kw) -> "ResamplerFir":
        return cls(channels, SampleRate(int(input_rate)), SampleRate(int(output_rate)), latency,
                   attenuation, **kw)

    @classmethod
    def new_from_hz(cls, channels, input_rate_hz: int, output_rate_hz: int,
                    latency: Latency = Latency.Sample64,
                    attenuation: Attenuation = Attenuation.Db120, **kw) -> "ResamplerFir":
        """Raises ValueError("input sample rate must be greater than zero") /
        ("output ...") where the reference panics (resampler_fir.rs:302-309)."""
        self = cls.__new__(cls)
        self._b = FirBatch(1, channels, input_rate_hz, output_rate_hz, latency, attenuation,
                           kw.get("device", 0), kw.get("kernel", Kernel.AUTO))
        self.channels = channels
        return self

    def buffer_size_output(self) -> int:
        return self._b.buffer_size_output()

    def resample(self, input: np.ndarray, output: np.ndarray):
        """-> (consumed, produced) in values; raises ResampleError (resampler_fir.rs:514-519)."""
        return self._b.resample(0, input, output)

    def delay(self) -> int:
        return self._b.delay()

    def reset(self) -> None:
        self._b.reset(0)

    def close(self) -> None:
        self._b.close()


# ---- host-only helpers (no GPU): used by the CPU test-suite ---------------------------------
def host_design_table(input_rate_hz: int, output_rate_hz: int, latency: Latency,
                      attenuation: Attenuation):
    lib = _lib.load()
    taps = Latency(latency).taps()
    out = np.empty((1024, taps), np.float32)
    bits = C.c_uint32(0)
    _check(lib.rsb_host_design_table(input_rate_hz, output_rate_hz, int(latency), int(attenuation),
                                     out.ctypes.data_as(_lib.f32p), out.size, C.byref(bits)))
    return out, bits.value


def device_design_table(input_rate_hz: int, output_rate_hz: int, latency: Latency,
                        attenuation: Attenuation, device: int = 0):
    """The same table designed ON the GPU (filter_design_device.cu) -> (table, device milliseconds)."""
    lib = _lib.load()
    taps = Latency(latency).taps()
    out = np.empty((1024, taps), np.float32)
    ms = C.c_float(0)
    _check(lib.rsb_device_design_table(device, input_rate_hz, output_rate_hz, int(latency), int(attenuation),
                                       out.ctypes.data_as(_lib.f32p), out.size, C.byref(ms)))
    return out, ms.value


def set_device_filter_design(enable: bool) -> bool:
    """Tables missing from the process-wide cache are designed on the GPU from now on; returns the
    previous setting."""
    return bool(_lib.load().rsb_set_device_filter_design(1 if enable else 0))


def host_plan(input_rate_hz: int, output_rate_hz: int, latency: Latency, position_bits: int,
              available: int, total_frames: int, call_frames: int, cap_frames: int,
              single_call: bool, max_calls: int = 1 << 16, max_frames: int = 1 << 22):
    """Runs the device's planner code on the host.  Returns dict (frames, not values)."""
    lib = _lib.load()
    pb = C.c_uint64(position_bits)
    av = C.c_uint32(available)
    cc = np.zeros(max_calls, np.uint32)
    cp = np.zeros(max_calls, np.uint32)
    arrs = [np.zeros(max_frames, np.uint32) for _ in range(4)]
    nf, ns = C.c_size_t(0), C.c_size_t(0)
    n_calls = lib.rsb_host_plan(input_rate_hz, output_rate_hz, int(latency), C.byref(pb),
                                C.byref(av), total_frames, call_frames, cap_frames,
                                int(single_call), cc.ctypes.data_as(_lib.u32p),
                                cp.ctypes.data_as(_lib.u32p), max_calls,
                                *[a.ctypes.data_as(_lib.u32p) for a in arrs], max_frames,
                                C.byref(nf), C.byref(ns))
    k = min(nf.value, max_frames)
    m = min(n_calls, max_calls)
    return {"calls": n_calls, "consumed": cc[:m], "produced": cp[:m], "n_frames": nf.value,
            "n_segments": ns.value, "input_offset": arrs[0][:k], "phase1": arrs[1][:k],
            "phase2": arrs[2][:k], "frac_bits": arrs[3][:k], "position_bits": pb.value,
            "available": av.value}
