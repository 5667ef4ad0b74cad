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

    # demod fixture: the reference-native shape (u8 IQ at 240 kS/s, src/demod.rs:70-117) -> 48 kHz baseband + power
    st = tx.control_channel(950, 3, lead_idle=10)
    iq = tx.iq_to_u8(tx.modulate_iq(st.dibits, 240_000, snr_db=25, cfo_hz=120.0, seed=5, amplitude=0.35))[: 2 * 40_000]
    chain = po.DemodChain(po.FMT_U8, False)
    parts = [chain.feed(iq[2 * a: 2 * b], want_power=True) for a, b in ((0, 16384), (16384, 32768), (32768, 40_000))]
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "demod_golden.npz"), iq_u8=iq,
                        baseband=np.concatenate([p[0] for p in parts]), power_dbm=np.array([p[1] for p in parts], dtype=np.float32))

    # channelizer fixture: 19.2 MS/s capture -> channel spectra after the channel-select filter (oracle/pfb_oracle.py:
    # channel_filter(channelize(x)), what the kernels deliver), the last 8 output times of 96 channels
    from oracle import pfb_oracle as pfb
    st = tx.control_channel(951, 1, lead_idle=0)
    cap = tx.wideband_capture({5: (st.dibits, 0.05, 0.0), 1530: (st.dibits[::-1].copy(), 0.03, 50.0)}, 12_000, noise_db=-50.0, seed=9)
    y = pfb.channel_filter(pfb.channelize(cap))
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "pfb_golden.npz"), capture=cap, rows=np.arange(22, 30),
                        channels=np.arange(0, 1536, 16) + 5 % 16, spectra=y[22:30][:, (np.arange(0, 1536, 16) + 5 % 16)].astype(np.complex64))
    print("demod", sum(len(p[0]) for p in parts), "baseband samples; pfb", y.shape)


if __name__ == "__main__":
    main()
