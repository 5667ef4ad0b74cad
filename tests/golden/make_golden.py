#!/usr/bin/env python3
"""Regenerates tests/golden/decode_golden.npz.

The reference holds no vector for this path and cannot run here, so the
committed fixture pins the build's own oracle: seeded transmitter output -> oracle events.
It guards the GPU path on the B200 box (where oracle and GPU are both checked against it) and
guards the oracle itself against silent drift (tests/test_golden_cpu.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "spec")]
from oracle import pyoracle as po  # noqa: E402
from tools import p25tx as tx  # noqa: E402


def main():
    rows = []
    for s in range(4):
        st = tx.control_channel(900 + s, 2, lead_idle=25 + s) if s < 2 else tx.traffic_channel(900 + s, 1, lead_idle=30)
        bb, _ = tx.baseband_48k(st.dibits, snr_db=[None, 14, 20, 10][s], dc=0.01 * s, seed=s, timing_offset=2.5 * s)
        rows.append(bb)
    n = min(len(r) for r in rows)
    # float16 rounding keeps the fixture small; the events are computed from the rounded samples
    bb = np.stack([r[:n] for r in rows]).astype(np.float16).astype(np.float32)
    ev = np.concatenate([po.MessageReceiver(stream=s).feed(bb[s]) for s in range(len(bb))])
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "decode_golden.npz"), baseband=bb.astype(np.float16),
                        events=ev.view(np.uint8))
    print(len(ev), "events")


if __name__ == "__main__":
    main()
