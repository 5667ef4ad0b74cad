"""Shared helpers for the parity tests."""
from __future__ import annotations

import numpy as np


def events_key(ev: np.ndarray):
    """Comparable view of an event array: (stream, sample, kind, payload bytes)."""
    return [(int(e["stream"]), int(e["sample"]), int(e["kind"]), bytes(e["payload"][: int(e["len"])])) for e in ev]


def oracle_events(po, streams_bb: np.ndarray, chunk: int | None = None):
    """Run one oracle MessageReceiver per row, return events ordered by (stream, sample) + stats."""
    out, stats = [], []
    for s, bb in enumerate(streams_bb):
        rx = po.MessageReceiver(stream=s)
        if chunk is None:
            out.append(rx.feed(bb))
        else:
            for i in range(0, len(bb), chunk):
                out.append(rx.feed(bb[i:i + chunk]))
        stats.append(rx.stats())
    return np.concatenate(out), np.stack(stats)


def check_against_truth(ev: np.ndarray, truth, stream: int = 0):
    """Events of one stream must equal the transmitter's ground truth (kind, payload) in order."""
    got = [(int(e["kind"]), bytes(e["payload"][: int(e["len"])])) for e in ev if int(e["stream"]) == stream]
    assert len(got) == len(truth), (len(got), len(truth))
    for i, (g, t) in enumerate(zip(got, truth)):
        if g[0] == 6 and t[0] == 6:      # VoiceFrame: compare u0..u7; the error counts depend on the channel
            g, t = (g[0], g[1][:32]), (t[0], t[1][:32])
        assert g == t, (i, g, t)
