run() { python bench.py --steps 4 --warmup 2 --no-cpu --no-e2e --seconds 30 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['roofline']['conv_ms_per_launch'], d['ms_per_step'])"; }
echo -n "base: "; run
echo -n "in stride 16K: "; RSB_DEBUG_FAKE_IN_STRIDE=16384 run
echo -n "out stride 16K: "; RSB_DEBUG_FAKE_OUT_STRIDE=16384 run
echo -n "both 16K: "; RSB_DEBUG_FAKE_IN_STRIDE=16384 RSB_DEBUG_FAKE_OUT_STRIDE=16384 run
echo -n "both 2M+4K: "; RSB_DEBUG_FAKE_IN_STRIDE=2101248 RSB_DEBUG_FAKE_OUT_STRIDE=2101248 run
echo -n "both 256K: "; RSB_DEBUG_FAKE_IN_STRIDE=262144 RSB_DEBUG_FAKE_OUT_STRIDE=262144 run
echo -n "128 streams x 240 s: "; python bench.py --steps 4 --warmup 2 --no-cpu --no-e2e --seconds 240 --streams 128 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['roofline']['conv_ms_per_launch'], d['ms_per_step'])"
echo -n "4096 streams x 7.5 s: "; python bench.py --steps 4 --warmup 2 --no-cpu --no-e2e --seconds 7.5 --streams 4096 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['roofline']['conv_ms_per_launch'], d['ms_per_step'])"
