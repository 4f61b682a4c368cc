source tools/experiments/run_fn.sh
for cfg in "0 0" "7 0" "7 1" "0 0" "7 1" "7 0"; do
  set -- $cfg; echo -n "variant=$1 split=$2: "; RSB_TC_VARIANT=$1 RSB_TC_EPI_SPLIT=$2 run; done
echo "--- power-capped (40 steps)"
export STEPS=40 WARM=10
for cfg in "0 0" "7 0" "7 1"; do
  set -- $cfg; echo -n "variant=$1 split=$2: "; RSB_TC_VARIANT=$1 RSB_TC_EPI_SPLIT=$2 run; done
