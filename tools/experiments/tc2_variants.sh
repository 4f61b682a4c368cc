run() { timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --seconds ${SECS:-60} 2>&1 | python -c "import sys,json; l=sys.stdin.read().strip().splitlines()[-1]; 
try:
    d=json.loads(l); print(d['roofline']['conv_ms_per_launch'], d['ms_per_step'], d['clocks']['sm_mhz'])
except Exception: print(l[:600])"; }
for v in ${VARS:-0 5 1 4 0 5}; do echo -n "variant=$v: "; RSB_TC_VARIANT=$v run; done
echo "--- skeleton (ablate 63), 30 s"
for v in 0 1; do echo -n "variant=$v: "; SECS=30 RSB_TC_ABLATE=63 RSB_TC_VARIANT=$v run; done
echo "--- v1 kernel"; echo -n "v1: "; RSB_TC_V1=1 run
