run() { python bench.py --steps 4 --warmup 2 --no-cpu --no-e2e --seconds 30 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['roofline']['conv_ms_per_launch'], d['ms_per_step'])"; }
echo -n "default (2 epi teams): "; run
echo -n "1 epi team: "; RSB_TC_EPI_TEAMS=1 run
echo -n "ablate=7 (floor): "; RSB_TC_ABLATE=7 run
echo -n "ablate=7, 1 team: "; RSB_TC_EPI_TEAMS=1 RSB_TC_ABLATE=7 run
echo -n "ablate=2 (no in): "; RSB_TC_ABLATE=2 run
echo -n "ablate=4 (no out): "; RSB_TC_ABLATE=4 run
echo -n "run_tiles 128: "; RSB_TC_RUN_TILES=128 run
echo -n "issuers 1: "; RSB_TC_ISSUERS=1 run
