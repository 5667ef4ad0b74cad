#!/bin/bash
# A/B of the u8 /5 warp kernel (P25CU_DDC5=3) against the variant with the decimator on the integer tensor pipe (=11):
# parity subset, worst baseband error vs the oracle, serial kernel times at the cfg5 and cfg4 shapes.  Writes gpurun_out/.
mkdir -p gpurun_out
{
for v in 3 11; do
  echo "== P25CU_DDC5=$v"
  P25CU_DDC5=$v python tests/ab/imma_err.py 2>&1 | tail -1
done
echo "== parity subset with P25CU_DDC5=11"
P25CU_DDC5=11 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "demod or process or cfg5 or cfg4 or golden or full_size" 2>&1 | tail -5
for v in 3 11; do
  echo "== shape_bench u8 /5 65536 control, P25CU_DDC5=$v"
  P25CU_DDC5=$v python tools/shape_bench.py --fmt u8 --decim 5 --streams 65536 2>/dev/null | tail -1
done
for v in 3 11; do
  echo "== shape_bench u8 /5 16384 traffic, P25CU_DDC5=$v"
  P25CU_DDC5=$v python tools/shape_bench.py --fmt u8 --decim 5 --streams 16384 --kind traffic 2>/dev/null | tail -1
done
} > gpurun_out/r02_imma_ab.txt 2>&1
cat gpurun_out/r02_imma_ab.txt
