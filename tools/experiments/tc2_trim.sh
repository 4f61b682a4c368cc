source tools/experiments/run_fn.sh
for a in 0 64 0 64; do echo -n "full clocks, ablate=$a: "; RSB_TC_ABLATE=$a run; done
export STEPS=40 WARM=10
for a in 0 64 0 64; do echo -n "capped, ablate=$a: "; RSB_TC_ABLATE=$a run; done
