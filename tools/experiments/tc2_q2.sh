source tools/experiments/run_fn.sh
legs() { python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('head', d['roofline']['conv_ms_per_launch'], d['roofline']['frac'], {k:(v.get('roofline',{}).get('conv_ms_per_launch'), v.get('roofline',{}).get('frac')) for k,v in d['other_configs'].items()})"; }
echo -n "default: "; legs
echo -n "split=0: "; RSB_TC_EPI_SPLIT=0 legs
echo -n "split=1: "; RSB_TC_EPI_SPLIT=1 legs
