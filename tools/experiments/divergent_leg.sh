# configs[2] (ii) leg of bench.py alone, plus the uniform small-submit shapes
python - <<'PY'
import sys, json
sys.path.insert(0, ".")
import bench
from resampler_b200 import _lib
for _ in range(2):
    r = bench.divergent_leg(_lib.load(), 0, 0, None)
    print(json.dumps({k: r[k] for k in r if k in ("value", "unit", "us_per_submit", "ms", "submits")} or r))
PY
python tools/stream_calls.py 4096 1 16000 48000 1 160 2>&1 | tail -1
python tools/stream_calls.py 1024 2 44100 48000 3 512 exact 2>&1 | tail -1
python tools/stream_calls.py 512 8 96000 48000 2 512 2>&1 | tail -1
python -m pytest tests/test_gpu_parity.py -q -x -k "submit or fused or divergent or async or handoff" 2>&1 | tail -15
