"""The committed golden fixture pins the oracle (CPU)."""
import os

import numpy as np
from util import events_key


def test_oracle_reproduces_golden(oracle):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "decode_golden.npz"))
    bb = g["baseband"].astype(np.float32)
    exp = g["events"].view(oracle.EVENT_DTYPE).reshape(-1)
    ev = np.concatenate([oracle.MessageReceiver(stream=s).feed(bb[s]) for s in range(len(bb))])
    assert events_key(ev) == events_key(exp)
    assert len(exp) > 20 and set(exp["kind"].tolist()) >= {1, 6, 7}


def test_oracle_reproduces_demod_golden(oracle):
    """u8 IQ at 240 kS/s in the reference's chunk sizes -> baseband + power (src/demod.rs:70-117, :123-134)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "demod_golden.npz"))
    iq, exp, pw = g["iq_u8"], g["baseband"], g["power_dbm"]
    chain = oracle.DemodChain(oracle.FMT_U8, False)
    parts = [chain.feed(iq[2 * a: 2 * b], want_power=True) for a, b in ((0, 16384), (16384, 32768), (32768, 40_000))]
    got = np.concatenate([p[0] for p in parts])
    assert len(got) == len(exp) == 8000
    assert np.max(np.abs(got - exp)) < 1e-6                  # same source, same flags: libm's atan2f is the only slack
    assert np.max(np.abs(np.array([p[1] for p in parts]) - pw)) < 1e-4
    assert exp.std() > 0.1                                   # a real C4FM signal, not silence


def test_channelizer_oracle_reproduces_golden():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
    from oracle import pfb_oracle as pfb
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "pfb_golden.npz"))
    y = pfb.channel_filter(pfb.channelize(g["capture"]))
    got = y[g["rows"]][:, g["channels"]]
    assert np.max(np.abs(got - g["spectra"])) < 1e-6 * np.max(np.abs(g["spectra"])) + 1e-9
    k5 = list(g["channels"]).index(5)
    assert np.abs(g["spectra"][:, k5]).mean() > 20 * np.median(np.abs(g["spectra"]))   # the carrier sits in channel 5
