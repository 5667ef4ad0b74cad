#!/usr/bin/env python3
"""shape_bench.py -- per-kernel device timing of the hot path on other shapes than bench.py's headline
(BASELINE.json configs[0], [3], [4]: the reference-native 240 kS/s u8 chain, voice traffic, 65,536 streams).

    python tools/shape_bench.py --fmt u8 --decim 5 --streams 65536 --kind control --steps 10

Input is one period (150 ms control / one HDU+LDU pair+TDULC superframe for voice) of a phase-continuous
synthetic transmission per base seed, tiled over the streams with per-stream circular shifts ON THE DEVICE
(SURVEY.md section 8d cfg5), so consecutive steps are a continuous signal and every step decodes real units.
Prints one JSON line: serialised ddc_fm / walker times, overlapped step time, Msamples/s, channels, roofline.
"""
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


def base_streams(kind: str, fs: int, n_base: int, snr_db: float):
    from tools import p25tx as tx
    out = []
    for b in range(n_base):
        st = tx.control_channel(1000 + b, 2, lead_idle=0) if kind == "control" else tx.traffic_channel(2000 + b, 1, lead_idle=0)
        out.append(tx.modulate_iq_periodic(st.dibits, fs, snr_db=snr_db, cfo_cycles=3 * (b - n_base // 2), seed=b))
    return np.stack(out)


def tile_on_device(base_dev, S: int, first: int = 0):
    """dev[s] = roll(base[(first+s) % B], -shift(first+s)) built in slabs to bound index memory."""
    import torch
    B, n = base_dev.shape[0], base_dev.shape[1]
    dev = torch.empty((S,) + tuple(base_dev.shape[1:]), dtype=base_dev.dtype, device=base_dev.device)
    ar = torch.arange(n, device=base_dev.device)
    slab = max(1, (1 << 27) // n)
    for s0 in range(0, S, slab):
        g = torch.arange(first + s0, first + min(S, s0 + slab), device=base_dev.device)
        sh = (g // B) * 5003 % n
        idx = (ar[None, :] + sh[:, None]) % n
        dev[s0:s0 + len(g)] = base_dev[(g % B)[:, None], idx]
    return dev


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fmt", default="u8", choices=["u8", "cf32"])
    ap.add_argument("--decim", type=int, default=5)
    ap.add_argument("--streams", type=int, default=65536)
    ap.add_argument("--kind", default="control", choices=["control", "traffic"])
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--snr", type=float, default=20.0)
    ap.add_argument("--n-base", type=int, default=16)
    ap.add_argument("--periods", type=int, default=1, help="periods of the transmission per step (longer chunks per call)")
    args = ap.parse_args()

    import torch
    import p25rx_b200 as p25
    from tools import p25tx as tx

    fs = 48000 * args.decim
    base = base_streams(args.kind, fs, args.n_base, args.snr)
    if args.periods > 1:
        base = np.tile(base, (1, args.periods))
    n = base.shape[1]
    if args.fmt == "u8":
        b8 = np.stack([tx.iq_to_u8(b).reshape(n, 2) for b in base])
        base_dev = torch.from_numpy(b8).cuda()
        fmt, bps = p25.FMT_U8_IQ, 2
    else:
        base_dev = torch.from_numpy(base.view(np.float32).reshape(len(base), n, 2)).cuda()
        fmt, bps = p25.FMT_CF32_IQ, 8
    S, K, W = args.streams, args.steps, args.warmup
    dev = tile_on_device(base_dev, S)
    torch.cuda.synchronize()          # the library runs on its own stream: the input must be complete before it is handed over
    ctx = p25.Context(S, fmt=fmt, decimation=args.decim, max_chunk_samples=n, device=0,
                      event_slots=(32 * (K + W + 6) + 64) * args.periods)
    stream = torch.cuda.ExternalStream(ctx.cuda_stream, device=0)
    for _ in range(W):
        ctx.process(dev, n)
        ctx.poll(copy=False)
    ctx.sync()
    # overlapped (product) timing: process + drain of the step's events, every step
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        t0.record(stream)
        for _ in range(K):
            ctx.process(dev, n)
            ctx.poll(copy=False)
        t1.record(stream)
    ctx.sync()
    step_ms = t0.elapsed_time(t1) / K
    ctx.process(dev, n)                      # one more, untimed, step for the event statistics
    ev = ctx.poll(copy=True)
    K_ev = 1
    # serialised per-kernel timing
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
    ddc_ms = sum(e[0].elapsed_time(e[1]) for e in bk) / len(bk)
    walk_ms = sum(e[1].elapsed_time(e[2]) for e in bk) / len(bk)
    ctx.poll(copy=False)
    kinds = np.bincount(ev["kind"], minlength=9)
    alg = S * n * (bps + 4.0 / args.decim)
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    line = {"shape": {"fmt": args.fmt, "decim": args.decim, "streams": S, "kind": args.kind, "samples_per_stream_per_step": n,
                      "input_gb_per_step": S * n * bps / 1e9, "snr_db": args.snr},
            "step_ms": step_ms, "ddc_ms_serial": ddc_ms, "walk_ms_serial": walk_ms,
            "msamples_per_s": S * n / (step_ms * 1e-3) / 1e6,
            "realtime_channels": S * n / args.decim / (step_ms * 1e-3) / 48000.0,
            "ddc_roofline": {"alg_bytes": alg, "achieved_gbs": alg / (ddc_ms * 1e-3) / 1e9, "peak": peak,
                             "frac": alg / (ddc_ms * 1e-3) / 1e9 / peak},
            "events_per_stream_per_step": len(ev) / S / K_ev,
            "event_kinds": {p25.EVENT_NAMES[i]: int(k) for i, k in enumerate(kinds) if k}}
    print(json.dumps(line))
    ctx.close()


if __name__ == "__main__":
    main()
