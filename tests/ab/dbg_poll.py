import sys, os, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'spec'); sys.path.insert(0, 'tests')
import p25rx_b200 as p25
from oracle import pyoracle as oracle
from tools import p25tx as tx
from util import events_key, oracle_events
S_ = 24
rows = []
for s in range(S_):
    st = tx.traffic_channel(6000 + s, 1) if s % 3 == 0 else tx.control_channel(6000 + s, 5)
    rows.append(tx.baseband_48k(st.dibits, snr_db=12 + s % 9, seed=s)[0])
n = min(len(r) for r in rows) // 4096 * 4096
bb = np.stack([r[:n] for r in rows])
ref, _ = oracle_events(oracle, bb)
def run(mode):
    a = p25.Context(S_, max_chunk_samples=1024, max_baseband=4096)
    got = []
    for k, i in enumerate(range(0, n, 4096)):
        a.decode(bb[:, i:i + 4096])
        if mode == 'records':
            got.append(a.poll())
        elif mode == 'packed_sync':
            w, ne, more = a.poll_packed(); got.append(a.unpack(w, ne))
        else:
            a.poll_start()
            if k:
                w, ne, more = a.poll_packed(); got.append(a.unpack(w, ne))
    if mode == 'packed_async':
        w, ne, more = a.poll_packed(); got.append(a.unpack(w, ne))
    a.close()
    ev = np.concatenate(got); return ev[np.lexsort((ev["sample"], ev["stream"]))]
kr = events_key(ref)
for mode in ('records', 'packed_sync', 'packed_async'):
    ev = run(mode); k = events_key(ev)
    bad = [(x, y) for x, y in zip(k, kr) if x != y]
    print(mode, len(k), len(kr), 'differing', len(bad))
    for x, y in bad[:4]:
        print('  gpu', x[:3], x[3].hex()); print('  ref', y[:3], y[3].hex())
