source tools/experiments/run_fn.sh
export STEPS=60 WARM=20 SECS=30
for a in 0 1 2 4 8 16 32 0; do echo -n "ablate=$a: "; RSB_TC_ABLATE=$a run; done
