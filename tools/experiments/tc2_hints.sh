source tools/experiments/run_fn.sh
echo "--- full 60 s"
for cfg in "0 10000000 10000000" "1 10000000 10000000" "0 0 10000000" "0 0 0" "0 1000 10000000" "0 1000 1000" "0 200 10000000" "0 10000000 10000000"; do
  set -- $cfg; echo -n "split=$1 crit=$2 other=$3: "; RSB_TC_EPI_SPLIT=$1 RSB_TC_HINT_CRIT=$2 RSB_TC_HINT_OTHER=$3 run; done
echo "--- skeleton 30 s"
for cfg in "0 10000000 10000000" "0 0 10000000" "0 0 0" "0 1000 1000"; do
  set -- $cfg; echo -n "split=$1 crit=$2 other=$3: "; SECS=30 RSB_TC_ABLATE=63 RSB_TC_EPI_SPLIT=$1 RSB_TC_HINT_CRIT=$2 RSB_TC_HINT_OTHER=$3 run; done
