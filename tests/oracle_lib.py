"""ctypes wrapper round the CPU oracle (oracle/fir_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, by ``__graft_entry__.smoke()``
and by ``bench.py``'s cpu_baseline / ``--impl reference`` legs.  The product
package ``resampler_b200`` never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
ORACLE_SO = ORACLE_DIR / "_build" / "liboracle.so"

_f32p = C.POINTER(C.c_float)
_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)
_szp = C.POINTER(C.c_size_t)


class _Trace(C.Structure):
    _fields_ = [
        ("input_offset", _u32p),
        ("phase1", _u32p),
        ("phase2", _u32p),
        ("frac_bits", _u32p),
        ("capacity", C.c_size_t),
        ("count", C.c_size_t),
    ]


def build_oracle(force: bool = False) -> Path:
    """Compile oracle/ with gcc (seconds).  Building the checker is not using it."""
    srcs = [ORACLE_DIR / n for n in ("fir_oracle.c", "fir_oracle.h", "cpu_bench.c", "Makefile")]
    stale = (not ORACLE_SO.exists()) or any(
        s.stat().st_mtime > ORACLE_SO.stat().st_mtime for s in srcs
    )
    if force or stale:
        env = dict(os.environ)
        env.setdefault("CC", "gcc")
        subprocess.run(["make", "-C", str(ORACLE_DIR), "-B"], check=True, env=env,
                       stdout=subprocess.DEVNULL)
    return ORACLE_SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(str(ORACLE_SO))
        L.orc_bessel_i0.restype = C.c_double
        L.orc_bessel_i0.argtypes = [C.c_double]
        L.orc_kaiser_window.argtypes = [C.c_size_t, C.c_double, C.c_int, _f32p]
        L.orc_cutoff_kaiser.restype = C.c_double
        L.orc_cutoff_kaiser.argtypes = [C.c_size_t, C.c_double]
        L.orc_make_sincs.argtypes = [C.c_size_t, C.c_size_t, C.c_float, C.c_double, C.c_int,
                                     _f32p, _f32p]
        L.orc_cutoff_for.restype = C.c_float
        L.orc_cutoff_for.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_double]
        L.orc_attenuation_beta.restype = C.c_double
        L.orc_attenuation_beta.argtypes = [C.c_int]
        L.orc_latency_taps.restype = C.c_int
        L.orc_latency_taps.argtypes = [C.c_int]
        for name in ("orc_convolve_avx512_order", "orc_convolve_scalar",
                     "orc_convolve_avx512_intrin"):
            fn = getattr(L, name)
            fn.restype = C.c_float
            fn.argtypes = [_f32p, _f32p, _f32p, C.c_float, C.c_size_t]
        L.orc_fir_new.restype = C.c_void_p
        L.orc_fir_new.argtypes = [C.c_size_t, C.c_uint32, C.c_uint32, C.c_int, C.c_int]
        L.orc_fir_free.argtypes = [C.c_void_p]
        L.orc_fir_set_conv.argtypes = [C.c_void_p, C.c_int]
        L.orc_fir_buffer_size_output.restype = C.c_size_t
        L.orc_fir_buffer_size_output.argtypes = [C.c_void_p]
        L.orc_fir_delay.restype = C.c_size_t
        L.orc_fir_delay.argtypes = [C.c_void_p]
        L.orc_fir_reset.argtypes = [C.c_void_p]
        L.orc_fir_resample_traced.restype = C.c_int
        L.orc_fir_resample_traced.argtypes = [C.c_void_p, _f32p, C.c_size_t, _f32p, C.c_size_t,
                                              _szp, _szp, C.POINTER(_Trace)]
        L.orc_fir_process.restype = C.c_size_t
        L.orc_fir_process.argtypes = [C.c_void_p, _f32p, C.c_size_t, C.c_size_t, C.c_size_t,
                                      _f32p, C.c_size_t, _szp, _szp, _u32p, _u32p, C.c_size_t,
                                      C.POINTER(_Trace)]
        L.orc_cpu_bench.restype = C.c_double
        L.orc_cpu_bench.argtypes = [C.c_size_t, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int,
                                    C.c_int, _f32p, C.c_size_t, C.c_size_t, C.c_size_t, _f32p,
                                    C.c_size_t, _u64p, C.c_int, C.c_int]
        L.orc_cpu_has_avx512f.restype = C.c_int
        L.orc_pcm_to_f32.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, _f32p]
        _lib = L
    return _lib


def _fp(a: np.ndarray):
    return a.ctypes.data_as(_f32p)


def _up(a: np.ndarray):
    return a.ctypes.data_as(_u32p)


LATENCY = {"Sample8": 0, "Sample16": 1, "Sample32": 2, "Sample64": 3}
ATTENUATION = {"Db60": 0, "Db90": 1, "Db120": 2}
CONV_AVX512, CONV_SCALAR, CONV_AVX512_INTRIN = 0, 1, 2


def bessel_i0(x: float) -> float:
    return lib().orc_bessel_i0(x)


def kaiser_window(n: int, beta: float, symmetric: bool) -> np.ndarray:
    out = np.empty(n, np.float32)
    lib().orc_kaiser_window(n, beta, int(symmetric), _fp(out))
    return out


def cutoff_kaiser(n: int, beta: float) -> float:
    return lib().orc_cutoff_kaiser(n, beta)


def make_sincs(sample_count: int, factor: int, cutoff: float, beta: float, symmetric: bool):
    out = np.empty((factor, sample_count), np.float32)
    s = C.c_float(0)
    lib().orc_make_sincs(sample_count, factor, np.float32(cutoff), beta, int(symmetric), _fp(out),
                         C.byref(s))
    return out, np.float32(s.value)


def cutoff_for(in_hz: int, out_hz: int, taps: int, beta: float) -> np.float32:
    return np.float32(lib().orc_cutoff_for(in_hz, out_hz, taps, beta))


def design_table(in_hz: int, out_hz: int, latency: int, attenuation: int) -> np.ndarray:
    """The [1024][taps] table the reference's constructor builds (resampler_fir.rs:406-421)."""
    L = lib()
    taps = L.orc_latency_taps(latency)
    beta = L.orc_attenuation_beta(attenuation)
    cutoff = cutoff_for(in_hz, out_hz, taps, beta)
    tab, _ = make_sincs(taps, 1024, cutoff, beta, True)
    return tab


def convolve(x, c1, c2, frac, kind=CONV_AVX512) -> np.float32:
    x = np.ascontiguousarray(x, np.float32)
    c1 = np.ascontiguousarray(c1, np.float32)
    c2 = np.ascontiguousarray(c2, np.float32)
    L = lib()
    fn = {CONV_AVX512: L.orc_convolve_avx512_order, CONV_SCALAR: L.orc_convolve_scalar,
          CONV_AVX512_INTRIN: L.orc_convolve_avx512_intrin}[kind]
    if kind == CONV_AVX512_INTRIN:
        # aligned loads of the coefficient rows (fir/avx512.rs:18-19)
        c1 = _aligned_copy(c1)
        c2 = _aligned_copy(c2)
    return np.float32(fn(_fp(x), _fp(c1), _fp(c2), np.float32(frac), len(x)))


def _aligned_copy(a: np.ndarray, align: int = 64) -> np.ndarray:
    raw = np.empty(a.nbytes + align, np.uint8)
    off = (-raw.ctypes.data) % align
    out = raw[off:off + a.nbytes].view(a.dtype).reshape(a.shape)
    out[...] = a
    return out


PCM_U8, PCM_S16, PCM_S24, PCM_S32, PCM_F32 = 0, 1, 2, 3, 4
PCM_BYTES = {PCM_U8: 1, PCM_S16: 2, PCM_S24: 3, PCM_S32: 4, PCM_F32: 4}


def pcm_to_f32(raw: np.ndarray, fmt: int, dup: int = 1) -> np.ndarray:
    """The CLI's format step (resample/src/main.rs:128-156): raw little-endian samples ->
    f32, each value written ``dup`` times in a row (2 = mono -> stereo)."""
    raw = np.ascontiguousarray(raw).view(np.uint8)
    n = raw.size // PCM_BYTES[fmt]
    out = np.empty(n * dup, np.float32)
    lib().orc_pcm_to_f32(raw.ctypes.data, fmt, n, dup, _fp(out))
    return out


class OracleFir:
    """Mirror of the reference's ``ResamplerFir`` over the C oracle."""

    def __init__(self, channels, in_hz, out_hz, latency=3, attenuation=2, conv=CONV_AVX512):
        self._h = lib().orc_fir_new(channels, in_hz, out_hz, latency, attenuation)
        if not self._h:
            raise ValueError("sample rate must be greater than zero")
        self.channels = channels
        lib().orc_fir_set_conv(self._h, conv)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_fir_free(self._h)
            self._h = None

    def buffer_size_output(self) -> int:
        return lib().orc_fir_buffer_size_output(self._h)

    def delay(self) -> int:
        return lib().orc_fir_delay(self._h)

    def reset(self) -> None:
        lib().orc_fir_reset(self._h)

    def resample(self, inp: np.ndarray, out: np.ndarray, trace: bool = False):
        """Returns (err, consumed, produced[, trace dict])."""
        inp = np.ascontiguousarray(inp, np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous
        c, p = C.c_size_t(0), C.c_size_t(0)
        tr = None
        arrays = None
        if trace:
            cap = max(1, len(out) // max(1, self.channels) + 2)
            arrays = [np.zeros(cap, np.uint32) for _ in range(4)]
            tr = _Trace(_up(arrays[0]), _up(arrays[1]), _up(arrays[2]), _up(arrays[3]), cap, 0)
        err = lib().orc_fir_resample_traced(self._h, _fp(inp), len(inp), _fp(out), len(out),
                                            C.byref(c), C.byref(p),
                                            C.byref(tr) if tr is not None else None)
        if trace:
            n = tr.count
            t = {"input_offset": arrays[0][:n], "phase1": arrays[1][:n], "phase2": arrays[2][:n],
                 "frac_bits": arrays[3][:n]}
            return err, c.value, p.value, t
        return err, c.value, p.value

    def process(self, inp: np.ndarray, call_len: int, out_cap_len: int = 0, trace: bool = False,
                out_capacity: int | None = None):
        """Canonical caller loop (resample/src/main.rs:226-254).  Returns dict."""
        inp = np.ascontiguousarray(inp, np.float32)
        if out_capacity is None:
            out_capacity = self._estimate_out(len(inp))
        out = np.zeros(out_capacity, np.float32)
        max_calls = len(inp) // max(1, call_len) + 64 + (4096 * self.channels) // max(1, call_len)
        if out_cap_len:
            max_calls += out_capacity // out_cap_len
        max_calls = max(max_calls, 16) * 4
        cc = np.zeros(max_calls, np.uint32)
        pc = np.zeros(max_calls, np.uint32)
        tot, itot = C.c_size_t(0), C.c_size_t(0)
        tr = None
        arrays = None
        if trace:
            cap = out_capacity // max(1, self.channels) + 8
            arrays = [np.zeros(cap, np.uint32) for _ in range(4)]
            tr = _Trace(_up(arrays[0]), _up(arrays[1]), _up(arrays[2]), _up(arrays[3]), cap, 0)
        ncalls = lib().orc_fir_process(self._h, _fp(inp), len(inp), call_len, out_cap_len,
                                       _fp(out), out_capacity, C.byref(tot), C.byref(itot),
                                       _up(cc), _up(pc), max_calls,
                                       C.byref(tr) if tr is not None else None)
        assert tot.value <= out_capacity, "oracle output capacity estimate too small"
        assert ncalls <= max_calls, "oracle call-count estimate too small"
        res = {"out": out[:tot.value], "consumed_total": itot.value, "calls": ncalls,
               "consumed": cc[:ncalls].copy(), "produced": pc[:ncalls].copy()}
        if trace:
            n = tr.count
            res["trace"] = {"input_offset": arrays[0][:n], "phase1": arrays[1][:n],
                            "phase2": arrays[2][:n], "frac_bits": arrays[3][:n]}
        return res

    def _estimate_out(self, in_len: int) -> int:
        # generous: frames / ratio + slack; ratio recovered from buffer_size_output is awkward,
        # so over-allocate using the worst plausible up-sampling factor seen in tests (x400).
        return int(in_len * self._up_factor() + 4 * self.buffer_size_output() + 64)

    def _up_factor(self) -> float:
        # buffer_size_output = (ceil((4096 - taps)/ratio) + 2) * channels  =>  1/ratio estimate
        taps = self.delay() * 2
        frames = self.buffer_size_output() / self.channels - 2
        return max(frames / (4096 - taps), 1e-3) * 1.01


def cpu_bench(n_streams, channels, in_hz, out_hz, latency, attenuation, inp: np.ndarray,
              frames: int, call_frames: int, n_threads: int, use_avx512: bool = True,
              keep_output: bool = False):
    """Times the threaded CPU baseline; inp is [n_streams, frames*channels] float32."""
    inp = np.ascontiguousarray(inp, np.float32)
    assert inp.shape == (n_streams, frames * channels)
    produced = np.zeros(n_streams, np.uint64)
    out = None
    out_stride = 0
    if keep_output:
        out_stride = int(frames * channels * out_hz / in_hz) + 64
        out = np.zeros((n_streams, out_stride), np.float32)
    secs = lib().orc_cpu_bench(n_streams, channels, in_hz, out_hz, latency, attenuation, _fp(inp),
                               inp.shape[1], frames, call_frames,
                               _fp(out) if out is not None else None, out_stride,
                               produced.ctypes.data_as(_u64p), n_threads, int(use_avx512))
    return secs, produced, out
