run() { timeout 100 python bench.py --steps 4 --warmup 2 --no-cpu --no-e2e --seconds 30 2>&1 | python -c "import sys,json; l=sys.stdin.read().strip().splitlines()[-1]; 
try:
    d=json.loads(l); print(d['roofline']['conv_ms_per_launch'], d['ms_per_step'])
except Exception: print(l[:600])"; }
for a in ${ABL:-0 7 63 16 4 2 56}; do echo -n "ablate=$a: "; RSB_TC_ABLATE=$a run; done
