# A/B of the walker's sync prefilter on one box: library default (prefilter on for channelizer contexts only, consulted after
# two empty search steps), forced on for every context (P25CU_WALK_PREFILTER=1), forced on and consulted at every step (build
# with -DP25_PREFILTER_QUIET=0), and off everywhere.
run() {  # label, env...
  local label=$1; shift
  for rep in 1 2; do
    env "$@" python tools/shape_bench.py --fmt u8 --decim 5 --streams 65536 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg5-shape walker ms', '$label', round(d['walk_ms_serial'],4), 'step', round(d['step_ms'],4))"
  done
  env "$@" python tools/shape_bench.py --fmt u8 --decim 5 --streams 16384 --kind traffic 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg4-shape walker ms', '$label', round(d['walk_ms_serial'],4), 'step', round(d['step_ms'],4))"
  env "$@" python tools/wide_bench.py --captures 8 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg3 (8 captures) walker ms', '$label', round(d['walk_ms_serial'],4), 'step', round(d['step_ms'],4))"
}
run "default" P25CU_DUMMY=1
run "prefilter forced on (after 2 empty steps)" P25CU_WALK_PREFILTER=1
run "prefilter forced on, every step" P25CU_WALK_PREFILTER=1 P25CU_LIB=$PWD/build/libp25cu_walk_q0.so
run "prefilter off" P25CU_WALK_PREFILTER=0
run "prefilter forced on, after 1 empty step" P25CU_WALK_PREFILTER=1 P25CU_LIB=$PWD/build/libp25cu_walk_q1.so
