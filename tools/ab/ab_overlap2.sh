# the same through bench.py's device-resident leg (K steps enqueued back to back, one drain): cfg5
for cfg in "0 0 5" "1 0 5" "1 0 4" "1 4 4" "1 8 4" "1 6 3" "1 12 3"; do
  set -- $cfg
  P25CU_OVERLAP=$1 P25CU_WALK_PERSIST=$2 P25CU_W5_CTAS=$3 python bench.py --workload cfg5 --no-cpu --sustained 0 --e2e-steps 1 --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg5 overlap=$1 walker_ctas_per_sm=$2 demod_ctas_per_sm=$3 ms_per_step', round(d['ms_per_step'],4), 'value', round(d['value']))"
done
