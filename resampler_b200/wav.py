"""Minimal RIFF/WAVE reader and f32 writer for the CLI batch path.

Host-side only (file parsing is not GPU work).  Mirrors what the reference CLI gets from
hound 3.5 (resample/src/main.rs:84-86, 128-137, 198-213): integer PCM of 8/16/24/32 bits or
IEEE float 32 in, stereo IEEE float 32 out.  The sample data is returned RAW, exactly as it
sits in the file, so that the conversion to f32 happens on the GPU (csrc/pcm_ingest.cu).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from pathlib import Path

import numpy as np

from .fir import PcmFormat

WAVE_FORMAT_PCM, WAVE_FORMAT_IEEE_FLOAT, WAVE_FORMAT_EXTENSIBLE = 1, 3, 0xFFFE


@dataclass
class WavData:
    sample_rate: int
    channels: int
    bits_per_sample: int
    fmt: PcmFormat
    raw: np.ndarray       # the data chunk, uint8, whole frames only

    @property
    def frames(self) -> int:
        return self.raw.size // (self.channels * self.fmt.bytes_per_sample())


def read_wav(path) -> WavData:
    b = Path(path).read_bytes()
    if len(b) < 12 or b[:4] != b"RIFF" or b[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file")
    pos, spec, data = 12, None, None
    while pos + 8 <= len(b):
        cid, size = b[pos:pos + 4], struct.unpack_from("<I", b, pos + 4)[0]
        body = b[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            if len(body) < 16:
                raise ValueError(f"{path}: fmt chunk of {len(body)} bytes (at least 16 needed)")
            tag, ch, rate, _, _, bits = struct.unpack_from("<HHIIHH", body, 0)
            if tag == WAVE_FORMAT_EXTENSIBLE:
                if len(body) < 26:
                    raise ValueError(f"{path}: truncated WAVE_FORMAT_EXTENSIBLE fmt chunk")
                valid_bits = struct.unpack_from("<H", body, 18)[0]
                tag = struct.unpack_from("<H", body, 24)[0]      # first word of the sub-format GUID
                if valid_bits not in (0, bits):
                    # e.g. 24 valid bits in 32-bit containers: the scale of the format step would
                    # be wrong; refuse rather than decode with the container width
                    raise ValueError(f"{path}: {valid_bits} valid bits in {bits}-bit containers "
                                     "are not supported")
            if ch == 0:
                raise ValueError(f"{path}: zero channels")
            spec = (tag, ch, rate, bits)
        elif cid == b"data":
            data = body
            break
        pos += 8 + size + (size & 1)
    if spec is None or data is None:
        raise ValueError(f"{path}: missing fmt or data chunk")
    tag, ch, rate, bits = spec
    if tag == WAVE_FORMAT_IEEE_FLOAT and bits == 32:
        fmt = PcmFormat.F32
    elif tag == WAVE_FORMAT_PCM and bits in (8, 16, 24, 32):
        fmt = {8: PcmFormat.U8, 16: PcmFormat.S16, 24: PcmFormat.S24, 32: PcmFormat.S32}[bits]
    else:
        raise ValueError(f"{path}: unsupported WAV encoding (format tag {tag}, {bits} bits)")
    frame_bytes = ch * fmt.bytes_per_sample()
    raw = np.frombuffer(data, np.uint8)
    if len(data) < size:
        raise ValueError(f"{path}: data chunk announces {size} bytes, the file holds {len(data)}")
    raw = raw[:raw.size - raw.size % frame_bytes].copy()     # hound likewise reads whole frames only
    return WavData(rate, ch, bits, fmt, raw)


def write_wav_f32(path, samples: np.ndarray, sample_rate: int, channels: int = 2) -> None:
    """IEEE float 32 WAV, as the CLI writes it (main.rs:198-213)."""
    samples = np.ascontiguousarray(samples, "<f4")
    data = samples.tobytes()
    fmt = struct.pack("<HHIIHH", WAVE_FORMAT_IEEE_FLOAT, channels, sample_rate,
                      sample_rate * channels * 4, channels * 4, 32)
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"data" + \
        struct.pack("<I", len(data)) + data
    Path(path).write_bytes(b"RIFF" + struct.pack("<I", len(body)) + body)


def write_wav_pcm(path, raw: np.ndarray, sample_rate: int, channels: int, bits: int) -> None:
    """Integer PCM WAV from raw little-endian sample bytes (test fixtures)."""
    data = np.ascontiguousarray(raw).view(np.uint8).tobytes()
    fmt = struct.pack("<HHIIHH", WAVE_FORMAT_PCM, channels, sample_rate,
                      sample_rate * channels * bits // 8, channels * bits // 8, bits)
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"data" + \
        struct.pack("<I", len(data)) + data + (b"\0" if len(data) & 1 else b"")
    Path(path).write_bytes(b"RIFF" + struct.pack("<I", len(body)) + body)
