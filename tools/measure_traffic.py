#!/usr/bin/env python
"""tools/measure_traffic.py -- DRAM bytes of ONE headline convolution launch (configs[1] at full
size), measured with ncu, written to profiles/conv_traffic.json together with a hash of the
kernel's sources (bench.py refuses a figure whose stamp does not match the sources it runs) and
the git commit.  Under gpurun:  python tools/measure_traffic.py [tensor|fast]"""
import csv
import json
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    kernel = sys.argv[1] if len(sys.argv) > 1 else "tensor"
    pattern = {"tensor": "conv_tc2", "fast": "conv_fast"}[kernel]
    out_csv = ROOT / "gpurun_out" / f"conv_{kernel}_dram_60s.csv"
    out_csv.parent.mkdir(exist_ok=True)
    cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum",
           "--clock-control", "none", "-k", f"regex:{pattern}", "-s", "1", "-c", "1", "--csv", "--log-file",
           str(out_csv), sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "1", "--no-e2e",
           "--no-cpu", "--no-legs", "--kernel", kernel]
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, timeout=600)
    rows = list(csv.reader(out_csv.open()))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    vals, name = {}, ""
    for r in rows[hdr + 1:]:
        d = dict(zip(rows[hdr], r))
        vals[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
        name = d["Kernel Name"]
    unit_scale = 1.0
    rd, wr = vals["dram__bytes_read.sum"] * unit_scale, vals["dram__bytes_write.sum"] * unit_scale
    # ncu may report Gbyte / Mbyte: normalise through the unit column
    for r in rows[hdr + 1:]:
        d = dict(zip(rows[hdr], r))
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(d["Metric Unit"])
        if scale and d["Metric Name"] == "dram__bytes_read.sum":
            rd = float(d["Metric Value"].replace(",", "")) * scale
        if scale and d["Metric Name"] == "dram__bytes_write.sum":
            wr = float(d["Metric Value"].replace(",", "")) * scale
    try:
        sha = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True,
                             text=True).stdout.strip() or "n/a (gpurun snapshot has no .git)"
    except Exception:
        sha = "n/a"
    alg = 4 * (1024 * 2 * 2879862 + 1024 * 2 * 2646000)
    tp = ROOT / "profiles" / "conv_traffic.json"
    allv = json.loads(tp.read_text()) if tp.exists() else {}
    allv[kernel] = {
        "kernel": name, "workload": "configs[1] 1024 stereo x 60 s, 44.1->48k, 128 taps (one launch per step)",
        "dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "dram_bytes_per_launch": int(rd + wr),
        "algorithmic_bytes_per_launch": alg, "ratio_to_algorithmic": round((rd + wr) / alg, 4),
        "kernel_ms_under_ncu": round(vals.get("gpu__time_duration.sum", 0.0) / 1e6, 3),
        "source_stamp": bench.source_stamp(), "git": sha, "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
        "source": f"tools/measure_traffic.py {kernel} (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum, one launch)",
    }
    tp.write_text(json.dumps(allv, indent=1) + "\n")
    (ROOT / "gpurun_out" / "conv_traffic.json").write_text(json.dumps(allv, indent=1) + "\n")
    print(json.dumps(allv[kernel]))


if __name__ == "__main__":
    main()
