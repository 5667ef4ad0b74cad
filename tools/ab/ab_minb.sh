for lib in "" build/libp25cu_mb6.so build/libp25cu_mb7.so; do
  for rep in 1 2; do
  P25CU_LIB=${lib:+$PWD/$lib} python tools/shape_bench.py --fmt u8 --decim 5 --streams 65536 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('u8 /5 65536 streams, lib=${lib:-default (5 CTAs per SM)} ddc ms', round(d['ddc_ms_serial'],4))"
  done
done
