source tools/experiments/run_fn.sh
for cfg in "0 0 0" "6 0 0" "6 1 0" "6 1 4" "6 0 4" "0 0 0"; do
  set -- $cfg; echo -n "variant=$1 split=$2 prefetch=$3: "; RSB_TC_VARIANT=$1 RSB_TC_EPI_SPLIT=$2 RSB_TC_PREFETCH=$3 run; done
