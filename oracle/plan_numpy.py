"""Second, independent restatement of ResamplerFir's *state machine only*
(no samples), in plain Python floats (IEEE f64) -- TEST INFRASTRUCTURE ONLY.

It exists because no reference test pins (consumed, produced) sequences or
phase indices: two independently written restatements
(this one and oracle/fir_oracle.c) must agree bit for bit.

Follows hasenbanck/resampler v0.5.1 src/resampler_fir.rs:
  :311-313 ratio, :456-465 buffer_size_output, :521-528 append,
  :542-590 output loop, :592-602 consume.
"""
from __future__ import annotations

import math
import struct

PHASES = 1024
INPUT_CAPACITY = 4096
BUFFER_SIZE = 2 * INPUT_CAPACITY
TAPS = {0: 16, 1: 32, 2: 64, 3: 128}


def f32_bits(x: float) -> int:
    """`x as f32` (round to nearest even) -> bit pattern."""
    return struct.unpack("<I", struct.pack("<f", x))[0]


class PlanFir:
    def __init__(self, in_hz: int, out_hz: int, latency: int):
        if in_hz == 0 or out_hz == 0:
            raise ValueError("sample rate must be greater than zero")
        self.ratio = float(in_hz) / float(out_hz)                 # :311-313
        self.taps = TAPS[latency]
        self.read_position = 0
        self.available = 0
        self.position = 0.0

    def buffer_size_output_frames(self) -> int:                   # :456-465
        return int(math.ceil((INPUT_CAPACITY - self.taps) / self.ratio)) + 2

    def reset(self):                                              # :638-642
        self.read_position = 0
        self.available = 0
        self.position = 0.0

    def call(self, in_frames: int, out_cap_frames: int):
        """One resample() call in frames.  Returns (frames_to_copy, outputs) where
        outputs is a list of (input_offset, phase1, phase2, frac_bits)."""
        write_position = self.read_position + self.available      # :524
        remaining = max(BUFFER_SIZE - write_position, 0)          # :525
        to_copy = min(in_frames, remaining, INPUT_CAPACITY - self.available)  # :526-528
        self.available += to_copy                                 # :538
        outs = []
        while True:
            off = int(math.floor(self.position))                  # :544
            if off + self.taps > self.available:                  # :549
                break
            if len(outs) >= out_cap_frames:                       # :553
                break
            fract = self.position - math.trunc(self.position)     # :557
            phase_f = min(fract * float(PHASES), float(PHASES - 1))  # :562
            phase1 = int(phase_f)                                 # :563
            phase2 = min(phase1 + 1, PHASES - 1)                  # :564
            frac_bits = f32_bits(phase_f - float(phase1))         # :565
            outs.append((off, phase1, phase2, frac_bits))
            self.position += self.ratio                           # :589
        consumed = min(int(math.floor(self.position)), self.available)  # :596
        self.read_position += consumed                            # :600
        self.available -= consumed                                # :601
        self.position -= float(consumed)                          # :602
        if self.read_position > INPUT_CAPACITY:                   # :605-615
            self.read_position = 0
        return to_copy, outs
