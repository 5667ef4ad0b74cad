#!/bin/bash
# A/B of the u8 /50 stream kernel (P25CU_DDC50=0, FFMA2) against the kernel with both decimating stages on the integer
# tensor pipe (=1): parity subset, serial kernel times at the configs[1] shape in u8.
mkdir -p gpurun_out
{
echo "== parity subset with P25CU_DDC50=1"
P25CU_DDC50=1 timeout 150 python -m pytest tests -m gpu -x -q -k "demod or process or cfg2 or golden or full_size or abi or chunk" 2>&1 | tail -5
for v in 0 1; do
  echo "== shape_bench u8 /50 1024 control, P25CU_DDC50=$v"
  P25CU_DDC50=$v timeout 100 python tools/shape_bench.py --fmt u8 --decim 50 --streams 1024 2>&1 | tail -1
done
} > gpurun_out/r02_imma50_ab.txt 2>&1
cat gpurun_out/r02_imma50_ab.txt
