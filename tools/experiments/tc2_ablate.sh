run() { python bench.py --steps 4 --warmup 2 --no-cpu --no-e2e --seconds 30 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['roofline']['conv_ms_per_launch'], d['ms_per_step'])"; }
for a in 0 1 2 4 3 5 6 7; do echo -n "ablate=$a: "; RSB_TC_ABLATE=$a run; done
for g in 2 3; do echo -n "gstages=$g: "; RSB_TC_GSTAGES=$g run; done
for r in 8 16 32 64 128; do echo -n "run_tiles=$r: "; RSB_TC_RUN_TILES=$r run; done
echo -n "issuers=1: "; RSB_TC_ISSUERS=1 run
