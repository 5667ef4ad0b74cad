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
