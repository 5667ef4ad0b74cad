#!/bin/bash
# Does the walker hide beside the IMMA-decimator /5 kernel?  The default (4 CTAs x 126 registers per SM) leaves no registers;
# with 3 CTAs per SM (+3 % alone) 17 k registers and ~100 KB of shared memory are free for a persistent walker grid.
run() {
  local label=$1; shift
  env "$@" timeout 100 python tools/shape_bench.py --fmt u8 --decim 5 --streams 65536 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$label: step ms', round(d['step_ms'],4), 'ddc', round(d['ddc_ms_serial'],4), 'walk', round(d['walk_ms_serial'],4))"
}
{
run "serial, 4 CTAs/SM (default)"
run "overlap, 3 demod CTAs/SM + 2 walker CTAs/SM" P25CU_OVERLAP=1 P25CU_W5_CTAS=3 P25CU_WALK_PERSIST=2
run "overlap, 3 demod CTAs/SM + 4 walker CTAs/SM" P25CU_OVERLAP=1 P25CU_W5_CTAS=3 P25CU_WALK_PERSIST=4
run "overlap, 2 demod CTAs/SM + 8 walker CTAs/SM" P25CU_OVERLAP=1 P25CU_W5_CTAS=2 P25CU_WALK_PERSIST=8
run "overlap, 3 demod CTAs/SM, walker grid not persistent" P25CU_OVERLAP=1 P25CU_W5_CTAS=3 P25CU_WALK_PERSIST=0
} > gpurun_out/r02_imma_overlap_ab.txt 2>&1
cat gpurun_out/r02_imma_overlap_ab.txt
