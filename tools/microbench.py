"""Runs the device micro-benchmarks (csrc/microbench.cu) and writes gpurun_out/microbench.json.
Usage (under gpurun): python tools/microbench.py"""
import ctypes as C
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from resampler_b200 import _lib  # noqa: E402

NAMES = {0: "ffma_scalar_peak", 1: "ffma2_packed_peak",
         10: "core K8 C4 map1D scalar", 11: "core K8 C4 map1D packed",
         12: "core K8 C4 map2D scalar", 13: "core K8 C4 map2D packed",
         14: "core K8 C2 map1D scalar", 15: "core K8 C2 map1D packed",
         16: "core K8 C2 map2D scalar", 17: "core K8 C2 map2D packed",
         18: "core K4 C4 map2D scalar", 19: "core K4 C4 map2D packed",
         20: "core K16 C2 map1D scalar", 21: "core K16 C2 map1D packed",
         22: "core K4 C8 map2D scalar", 23: "core K4 C8 map2D packed",
         24: "core K8 C8 map2D scalar", 25: "core K8 C8 map2D packed",
         26: "core K16 C4 map2D scalar", 27: "core K16 C4 map2D packed"}


def main():
    lib = _lib.load()
    rows = []
    for id_, name in NAMES.items():
        for arg in ((2, 4, 8) if id_ < 10 else (1, 2, 3)):
            r, a = C.c_double(0), C.c_double(0)
            rc = lib.rsb_microbench(0, id_, arg, C.byref(r), C.byref(a))
            rows.append({"id": id_, "name": name, "arg": arg, "rc": rc, "tflops": round(r.value, 2),
                         "aux": a.value})
            print(f"{id_:3d} {name:28s} arg={arg} rc={rc} {r.value:8.2f} TFLOP/s aux={a.value}",
                  flush=True)
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / "microbench.json").write_text(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
