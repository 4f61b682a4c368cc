run() { timeout 100 python bench.py --steps ${STEPS:-5} --warmup ${WARM:-3} --no-cpu --no-e2e --no-legs --seconds ${SECS:-60} 2>&1 | python -c "import sys,json; l=sys.stdin.read().strip().splitlines()[-1]; 
try:
    d=json.loads(l); print(d['roofline']['conv_ms_per_launch'], d['ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception: print(l[:600])"; }
