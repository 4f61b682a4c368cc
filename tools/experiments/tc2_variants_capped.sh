source tools/experiments/run_fn.sh
export STEPS=40 WARM=10
for v in 0 2 7 3 0; do echo -n "capped variant=$v: "; RSB_TC_VARIANT=$v run; done
