#!/usr/bin/env python3
"""wide_bench.py -- device timing of the wideband channelizer path (BASELINE.json configs[2]):
one 19.2 MS/s cf32 capture ("20 MHz" nominal), 64 of its 1,536 12.5 kHz slots carrying control channels, split by
the polyphase filter bank, every channel demodulated and decoded.

    python tools/wide_bench.py [--captures 1] [--chunk-ms 100] [--steps 10]

One step = one chunk of every capture through p25cu_process + p25cu_poll.  Prints one JSON line with per-kernel
times (CUDA events around the serialised kernels), the real-time factor and the HBM roofline of the two kernels
(algorithmic bytes per input sample: 8 in + 15.36 baseband out)."""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "spec")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--captures", type=int, default=1)
    ap.add_argument("--chunk-ms", type=int, default=150)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--occupied", type=int, default=64)
    args = ap.parse_args()
    import torch
    import p25rx_b200 as p25
    from tools import p25tx as tx

    fs = 19_200_000
    n = fs * args.chunk_ms // 1000
    assert n % 400 == 0 and args.chunk_ms % 150 == 0, "chunks are whole 150 ms periods so that the signal repeats seamlessly"
    # one 150 ms period per occupied slot (two TSDUs), phase-continuous when repeated
    base = np.zeros(fs * 150 // 1000, dtype=np.complex128)
    slots = [(37 * i + 5) % 1536 for i in range(args.occupied)]
    for i, k in enumerate(slots):
        st = tx.control_channel(3000 + i, 2, lead_idle=0)
        f = (k if k < 768 else k - 1536) * 12500.0
        cyc = int(round(f * 0.150))                                     # whole carrier turns per period
        iq = tx.modulate_iq_periodic(st.dibits, fs, snr_db=None, cfo_cycles=cyc, seed=i, amplitude=0.01)
        base += np.roll(iq, 9973 * i)
    rng = np.random.default_rng(1)
    base += 10 ** (-60 / 20) / np.sqrt(2) * (rng.standard_normal(len(base)) + 1j * rng.standard_normal(len(base)))
    chunk = np.tile(base.astype(np.complex64), args.chunk_ms // 150)
    caps = np.stack([np.roll(chunk, 400 * 77 * c) for c in range(args.captures)])
    dev = torch.from_numpy(caps.view(np.float32).reshape(args.captures, n, 2)).cuda()
    torch.cuda.synchronize()
    S = 1536 * args.captures
    ctx = p25.Context(S, fmt=p25.FMT_CF32_IQ, decimation=400, max_chunk_samples=n, event_slots=64 * args.chunk_ms // 150 + 64)
    stream = torch.cuda.ExternalStream(ctx.cuda_stream, device=0)
    for _ in range(args.warmup):
        ctx.process(dev, n)
        ctx.poll(copy=False)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        t0.record(stream)
        for _ in range(args.steps):
            ctx.process(dev, n)
            ctx.poll(copy=False)
        t1.record(stream)
    ctx.sync()
    step_ms = t0.elapsed_time(t1) / args.steps
    ctx.set_overlap(False)
    bk = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(5)]
    with torch.cuda.stream(stream):
        for e3 in bk:
            e3[0].record(stream)
            ctx.demod(dev, n, want_baseband=False)
            e3[1].record(stream)
            ctx.decode()
            e3[2].record(stream)
    ctx.sync()
    demod_ms = sum(e[0].elapsed_time(e[1]) for e in bk) / len(bk)
    walk_ms = sum(e[1].elapsed_time(e[2]) for e in bk) / len(bk)
    ctx.poll(copy=False)
    ctx.process(dev, n)
    ev = ctx.poll()
    good = ev[ev["kind"] == p25.EV_TSBK]
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    alg = args.captures * n * (8 + 15.36)          # SURVEY 8d style: 8 B per input sample in, 4 B per channel and output time out
    line = {"shape": {"captures": args.captures, "sample_rate": fs, "chunk_ms": args.chunk_ms, "channels": S, "occupied": args.occupied,
                      "input_mb_per_step": args.captures * n * 8 / 1e6},
            "step_ms": step_ms, "channelizer_plus_baseband_ms_serial": demod_ms, "walk_ms_serial": walk_ms,
            "msamples_per_s": args.captures * n / (step_ms * 1e-3) / 1e6,
            "realtime_factor": args.chunk_ms / step_ms * args.captures,
            "roofline": {"alg_bytes": alg, "achieved_gbs": alg / (demod_ms * 1e-3) / 1e9, "peak": peak, "frac": alg / (demod_ms * 1e-3) / 1e9 / peak},
            "tsbk_per_step": int(len(good)), "tsbk_channels": int(len(set(int(s) % 1536 for s in good["stream"]))),
            "crc_ok": bool(all(__import__("p25_spec").crc_ccitt_p25(bytes(e["payload"][:10])) == (int(e["payload"][10]) << 8 | int(e["payload"][11]))
                               for e in good[:200]))}
    print(json.dumps(line))
    ctx.close()


if __name__ == "__main__":
    main()
