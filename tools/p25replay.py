#!/usr/bin/env python3
"""p25replay.py -- decode recordings on the GPU path and print what the reference's consumers would see.

    python tools/p25replay.py a.f32 b.f32 ...            `p25rx -r FILE` for many files at once: f32le / 48 kHz / mono
                                                         baseband (reference src/main.rs:95-98, src/replay.rs:26-57)
    python tools/p25replay.py --iq a.u8 b.u8 ...         raw RTL-SDR captures: interleaved u8 I/Q at 240 kS/s in
                                                         32,768-byte chunks (src/sdr.rs:25-33, src/demod.rs:62-119)

One JSON object per line: decoded events in (stream, sample) order per block (TSBKs with opcode / CRC like
RecvTask::handle_tsbk reads them, src/recv.rs:237-274), "sigPower" every 4th chunk in --iq mode (src/demod.rs:95-101),
and at the end every stream's stats in the hub's schema (src/hub.rs:557-581).  Replay ends with the shortest file."""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def event_json(p25, co, files, e):
    kind = p25.EVENT_NAMES[int(e["kind"])]
    pl = bytes(e["payload"][: int(e["len"])])
    d = {"file": files[int(e["stream"])], "sample": int(e["sample"]), "event": kind}
    if kind == "PacketNID":
        d.update(nac=pl[0] | pl[1] << 8, duid=pl[2])
    elif kind == "TrunkingControl":
        t = co.TsbkFields(pl)
        d.update(opcode=t.opcode() or f"0x{t.opcode_bits:02x}", mfg=t.mfg(), crc_valid=t.crc_valid(), payload=t.payload().hex())
    elif kind in ("LinkControl", "VoiceTerm"):
        lc = co.LinkControlFields(pl)
        d.update(opcode=lc.opcode() or f"0x{pl[0] & 0x3F:02x}", payload=lc.payload().hex())
    elif kind == "VoiceFrame":
        w = np.frombuffer(pl, dtype="<u4")
        d.update(chunks=[int(x) for x in w[:8]], errors=[int(x) for x in w[8:15]])
    elif kind in ("Error", "LowSpeedDataFragment"):
        d.update(value=int(np.frombuffer(pl[:4], dtype="<u4")[0]))
    else:
        d.update(payload=pl.hex())
    return d


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("files", nargs="+")
    ap.add_argument("--iq", action="store_true", help="files are raw u8 IQ at 240 kS/s instead of 48 kHz f32 baseband")
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args()
    import p25rx_b200 as p25
    from p25rx_b200 import consumers as co

    S = len(args.files)
    handles = [open(f, "rb") for f in args.files]
    out = sys.stdout
    if args.iq:
        ctx = p25.Context(S, fmt=p25.FMT_U8_IQ, decimation=5, max_chunk_samples=co.BUF_BYTES // 2, device=args.device)
        demod = p25.DemodTask(ctx)
        readers = [co.read_iq_chunks(h) for h in handles]
        for chunks in zip(*readers):
            _, power = demod.run_chunk(np.stack(chunks))
            if power is not None:
                out.write(json.dumps({"event": "sigPower", "dBm": [round(float(x), 2) if np.isfinite(x) else None for x in power]}) + "\n")
            ctx.decode()
            for e in ctx.poll():
                out.write(json.dumps(event_json(p25, co, args.files, e)) + "\n")
    else:
        rr = p25.ReplayReceiver(n_streams=S, device=args.device)
        ctx = rr.ctx
        for e in rr.replay(handles):
            out.write(json.dumps(event_json(p25, co, args.files, e)) + "\n")
    for s in range(S):
        out.write(json.dumps({"file": args.files[s], "event": "stats", "stats": co.stats_json(ctx.stats(s))}) + "\n")
    ctx.close()


if __name__ == "__main__":
    main()
