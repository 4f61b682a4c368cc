#!/usr/bin/env python
"""Generates tests/golden/fir_cases.npz -- committed regression fixtures of the CPU oracle.

The Rust reference cannot run in this image (no rustc / cargo), so these are NOT outputs of the
reference itself: they are outputs of oracle/fir_oracle.c (the line-by-line restatement, pinned
to the reference's own known-answer tests in tests/test_oracle_golden.py) on seeded inputs,
frozen so that (a) a later change of the oracle or of the host libm is noticed, and (b) the GPU
parity tests can compare the CUDA path against committed bytes.  Run from the repo root:

    python tests/golden/make_fixtures.py

Each case: one stream through the canonical caller loop (resample/src/main.rs:226-254).
Stored per case: the seeded input, per-call (consumed, produced), the per-frame plan
(input_offset, phase1, phase2, frac bits) and the output samples.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib as O  # noqa: E402

# name: channels, in_hz, out_hz, latency, attenuation, call frames, out-capacity frames, frames
CASES = {
    "c1_48k_44k1_t128": (2, 48000, 44100, 3, 1, 512, 0, 6000),     # BASELINE configs[0]
    "c2_44k1_48k_t128": (2, 44100, 48000, 3, 1, 512, 0, 6000),     # configs[1] / [4] parameters
    "c3_16k_48k_t32": (1, 16000, 48000, 1, 1, 160, 0, 4000),       # configs[2]
    "c4_96k_48k_t64_8ch": (8, 96000, 48000, 2, 1, 512, 0, 3000),   # configs[3]
    "cap100_44k1_48k": (2, 44100, 48000, 3, 1, 512, 100, 3000),    # capacity-limited calls
    "lookahead_384k_1k_t16": (1, 384000, 1000, 0, 1, 512, 0, 9000),  # ratio > taps (0.5.1 fix)
    "db120_24k_16k_t64": (2, 24000, 16000, 2, 2, 333, 0, 3000),    # arbitrary rates, Db120
}


def case_input(name: str, n: int) -> np.ndarray:
    seed = int.from_bytes(name.encode()[:8].ljust(8, b"\0"), "little") % (2 ** 32)
    return np.random.default_rng(seed).uniform(-1.0, 1.0, n).astype(np.float32)


def run_case(name):
    ch, in_hz, out_hz, lat, att, call, cap, frames = CASES[name]
    x = case_input(name, frames * ch)
    f = O.OracleFir(ch, in_hz, out_hz, lat, att)
    r = f.process(x, call * ch, out_cap_len=cap * ch, trace=True)
    return {"x": x, "out": r["out"], "consumed": r["consumed"], "produced": r["produced"],
            "input_offset": r["trace"]["input_offset"], "phase1": r["trace"]["phase1"],
            "phase2": r["trace"]["phase2"], "frac_bits": r["trace"]["frac_bits"]}


def main():
    arrays = {}
    for name in CASES:
        for k, v in run_case(name).items():
            arrays[f"{name}/{k}"] = v
    out = Path(__file__).resolve().parent / "fir_cases.npz"
    np.savez_compressed(out, **arrays)
    print(f"wrote {out} ({out.stat().st_size} bytes, {len(CASES)} cases)")


if __name__ == "__main__":
    main()
