"""ctypes binding of the C ABI declared in include/resampler_b200.h.

The shared library is built in-tree (``make`` / ``__graft_entry__.build()``) as
``resampler_b200/lib/libresampler_b200.so``.  There is no fallback of any kind:
if the library is missing, importing the compute classes raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "lib" / "libresampler_b200.so"

f32p = C.POINTER(C.c_float)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
szp = C.POINTER(C.c_size_t)
f32pp = C.POINTER(f32p)

# name -> (restype, argtypes); one entry per symbol include/resampler_b200.h declares
SIGNATURES = {
    "rsb_version": (C.c_char_p, []),
    "rsb_status_string": (C.c_char_p, [C.c_int]),
    "rsb_last_error": (C.c_char_p, []),
    "rsb_device_count": (C.c_int, []),
    "rsb_fir_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_uint32, C.c_uint32, C.c_uint32,
                                 C.c_uint32, C.c_int, C.c_int]),
    "rsb_fir_destroy": (None, [C.c_void_p]),
    "rsb_fir_set_kernel": (C.c_int, [C.c_void_p, C.c_int]),
    "rsb_fir_channels": (C.c_uint32, [C.c_void_p]),
    "rsb_fir_last_kernel": (C.c_int, [C.c_void_p]),
    "rsb_fir_n_streams": (C.c_uint32, [C.c_void_p]),
    "rsb_fir_taps": (C.c_uint32, [C.c_void_p]),
    "rsb_fir_ratio": (C.c_double, [C.c_void_p]),
    "rsb_fir_device": (C.c_int, [C.c_void_p]),
    "rsb_fir_buffer_size_output": (C.c_size_t, [C.c_void_p]),
    "rsb_fir_delay": (C.c_size_t, [C.c_void_p]),
    "rsb_fir_reset": (C.c_int, [C.c_void_p, C.c_int64]),
    "rsb_fir_coeffs": (C.c_int, [C.c_void_p, f32p, C.c_size_t]),
    "rsb_fir_resample": (C.c_int, [C.c_void_p, C.c_uint32, f32p, C.c_size_t, f32p, C.c_size_t,
                                   szp, szp]),
    "rsb_fir_submit_batch": (C.c_int, [C.c_void_p, C.c_uint32, u32p, C.POINTER(C.c_void_p), szp,
                                       C.POINTER(C.c_void_p), szp, szp, szp, C.c_int, C.c_uint32]),
    "rsb_fir_process_batch": (C.c_int, [C.c_void_p, C.c_uint32, u32p, C.POINTER(C.c_void_p), szp,
                                        C.c_size_t, C.c_size_t, C.POINTER(C.c_void_p), szp, szp,
                                        szp, u32p, C.c_int, C.c_uint32]),
    "rsb_fir_process_pcm_batch": (C.c_int, [C.c_void_p, C.c_uint32, u32p, C.POINTER(C.c_void_p),
                                            szp, C.c_int, C.c_uint32, C.c_size_t, C.c_size_t,
                                            C.POINTER(C.c_void_p), szp, szp, szp, u32p, C.c_int,
                                            C.c_uint32]),
    "rsb_fir_flush_batch": (C.c_int, [C.c_void_p, C.c_uint32, u32p, C.POINTER(C.c_void_p), szp, szp,
                                      C.c_int, C.c_uint32]),
    "rsb_fir_last_pcm_fused": (C.c_int, [C.c_void_p]),
    "rsb_fir_last_ingest_ms": (C.c_int, [C.c_void_p, f32p]),
    "rsb_fir_sync": (C.c_int, [C.c_void_p]),
    "rsb_fir_last_call_counts": (C.c_int, [C.c_void_p, C.c_uint32, u32p, u32p, C.c_size_t, szp]),
    "rsb_fir_last_plan": (C.c_int, [C.c_void_p, C.c_uint32, u32p, u32p, u32p, u32p, C.c_size_t,
                                    szp]),
    "rsb_fir_timer_start": (C.c_int, [C.c_void_p]),
    "rsb_fir_timer_stop": (C.c_int, [C.c_void_p, f32p]),
    "rsb_fir_conv_times": (C.c_int, [C.c_void_p, f32p, C.c_size_t, szp]),
    "rsb_debug_phase_cycles": (C.c_int, [C.c_void_p, C.c_int, u64p]),
    "rsb_debug_tc_hang": (C.c_int, [C.POINTER(C.c_uint32)]),
    "rsb_debug_tc_cycles": (C.c_int, [C.c_void_p, C.c_int, u64p, C.c_uint32]),
    "rsb_fir_launch_count": (C.c_uint64, [C.c_void_p]),
    "rsb_fir_cuda_stream": (C.c_void_p, [C.c_void_p]),
    "rsb_fir_host_pipeline_stats": (C.c_int, [C.c_void_p, u64p, u64p]),
    "rsb_fft_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]),
    "rsb_fft_destroy": (None, [C.c_void_p]),
    "rsb_fft_chunk_size_input": (C.c_size_t, [C.c_void_p]),
    "rsb_fft_chunk_size_output": (C.c_size_t, [C.c_void_p]),
    "rsb_fft_delay": (C.c_size_t, [C.c_void_p]),
    "rsb_fft_reset": (C.c_int, [C.c_void_p, C.c_int64]),
    "rsb_fft_resample": (C.c_int, [C.c_void_p, C.c_uint32, f32p, C.c_size_t, f32p, C.c_size_t]),
    "rsb_fft_process_batch": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_void_p),
                                        C.POINTER(C.c_size_t), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                        C.POINTER(C.c_size_t), C.c_int, C.c_uint32]),
    "rsb_fft_sync": (C.c_int, [C.c_void_p]),
    "rsb_fft_launch_count": (C.c_uint64, [C.c_void_p]),
    "rsb_fft_last_error": (C.c_char_p, []),
    "rsb_set_device_filter_design": (C.c_int, [C.c_int]),
    "rsb_host_sinf_restated": (C.c_float, [C.c_float]),
    "rsb_device_design_table": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_int, f32p, C.c_size_t,
                                          C.POINTER(C.c_float)]),
    "rsb_pcie_probe": (C.c_int, [C.c_int, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "rsb_alloc_pinned": (C.c_void_p, [C.c_size_t]),
    "rsb_free_pinned": (None, [C.c_void_p]),
    "rsb_alloc_device": (C.c_void_p, [C.c_int, C.c_size_t]),
    "rsb_free_device": (None, [C.c_int, C.c_void_p]),
    "rsb_memcpy": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]),
    "rsb_fill_synthetic": (C.c_int, [C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64,
                                     C.c_uint32, C.c_uint32, C.c_uint64]),
    "rsb_microbench": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double),
                                 C.POINTER(C.c_double)]),
    "rsb_host_bessel_i0": (C.c_double, [C.c_double]),
    "rsb_host_cutoff_kaiser": (C.c_double, [C.c_uint32, C.c_double]),
    "rsb_host_kaiser_window": (None, [C.c_uint32, C.c_double, C.c_int, f32p]),
    "rsb_host_make_sincs": (None, [C.c_uint32, C.c_uint32, C.c_float, C.c_double, C.c_int, f32p]),
    "rsb_host_design_table": (C.c_int, [C.c_uint32, C.c_uint32, C.c_int, C.c_int, f32p,
                                        C.c_size_t, u32p]),
    "rsb_host_plan": (C.c_size_t, [C.c_uint32, C.c_uint32, C.c_int, u64p, u32p, C.c_uint64,
                                   C.c_uint32, C.c_uint32, C.c_int, u32p, u32p, C.c_size_t, u32p,
                                   u32p, u32p, u32p, C.c_size_t, szp, szp]),
}

_lib = None


def load() -> C.CDLL:
    """Loads the native library; raises when it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make` (or __graft_entry__.build()); "
                "resampler_b200 has no CPU or pure-Python fallback")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)   # AttributeError if the .so does not export the symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    return load().rsb_last_error().decode()
