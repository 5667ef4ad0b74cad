#!/usr/bin/env python3
"""Worst |baseband - oracle| of the u8 /5 fast path for the variant selected by P25CU_DDC5 (3: FFMA2 warp kernel,
11: decimator on the integer tensor pipe), on carriers from +-4 LSB to full scale and on four chunk phases."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # tests/ab/ -> repository root
for p in (ROOT, os.path.join(ROOT, "spec")):
    sys.path.insert(0, p)
import p25rx_b200 as p25           # noqa: E402
from oracle import pyoracle as oracle   # noqa: E402  (test tooling: the checker, never the product path)
from tools import p25tx as tx      # noqa: E402

st = tx.control_channel(321, 3)
amps = [0.01, 0.03, 0.06, 0.5, 0.95]
rows = [tx.iq_to_u8(tx.modulate_iq(st.dibits, 240_000, snr_db=25, cfo_hz=200.0 * (s - 1.5), seed=s, amplitude=a))
        for s, a in enumerate(amps)]
n = min(len(r) for r in rows) // 2
chunk = 16384
data = np.stack([r[: 2 * n] for r in rows])
ctx = p25.Context(len(amps), fmt=p25.FMT_U8_IQ, decimation=5, max_chunk_samples=chunk)
chains = [oracle.DemodChain(oracle.FMT_U8, False) for _ in amps]
worst = [0.0] * len(amps)
pw_worst = 0.0
for i in range(0, n - chunk + 1, chunk):
    part = np.ascontiguousarray(data[:, 2 * i: 2 * (i + chunk)])
    bb, n_out, pw = ctx.demod(part, chunk, want_power=True)
    for s in range(len(amps)):
        ref, pref = chains[s].feed(part[s], want_power=True)
        worst[s] = max(worst[s], float(np.max(np.abs(bb[s] - ref))))
        pw_worst = max(pw_worst, abs(pw[s] - pref))
print(json.dumps({"variant": os.environ.get("P25CU_DDC5", "default"), "amplitudes": amps, "worst_abs_err": worst,
                  "worst_power_db_err": float(pw_worst)}))
