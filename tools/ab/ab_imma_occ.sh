#!/bin/bash
# occupancy of the IMMA-decimator kernel: 3 / 4 CTAs per SM of the default build (118 registers), 5 CTAs of a build with
# __launch_bounds__(128, 5) (96 registers)
run() { # label, env...
  local label=$1; shift
  for rep in 1 2; do
    env P25CU_DDC5=11 "$@" python tools/shape_bench.py --fmt u8 --decim 5 --streams 65536 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$label: ddc ms', round(d['ddc_ms_serial'],4))"
  done
}
{
run "3 CTAs/SM, 118 regs" P25CU_W5_CTAS=3
run "4 CTAs/SM, 118 regs (default)"
run "5 CTAs/SM, 96 regs" P25CU_LIB=$PWD/build/libp25cu_w5i_m5.so
} > gpurun_out/r02_imma_occupancy_ab.txt 2>&1
cat gpurun_out/r02_imma_occupancy_ab.txt
