source tools/experiments/run_fn.sh
echo "--- full 60 s: variant x epi_split"
for cfg in "0 0" "0 1" "2 0" "2 1" "1 0" "3 0" "5 0" "0 0"; do
  set -- $cfg; echo -n "variant=$1 split=$2: "; RSB_TC_VARIANT=$1 RSB_TC_EPI_SPLIT=$2 run; done
echo "--- ablations, 30 s, variant 0"
for a in 0 1 2 4 8 16 32 6 7 63; do echo -n "ablate=$a: "; SECS=30 RSB_TC_ABLATE=$a run; done
echo "--- issuers=1"; echo -n "issuers=1: "; RSB_TC_ISSUERS=1 run
for r in 48 64 96 128 160; do echo -n "run_tiles=$r: "; RSB_TC_RUN_TILES=$r run; done
