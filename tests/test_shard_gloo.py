"""N > 1 host logic on CPU: world_size-2 gloo run of the stream sharding + event gather.

Each rank decodes its block of streams (with the oracle standing in for the GPU, this is a host-logic test),
rank 0 gathers; the result must equal the single-process decode of all streams."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _streams(n):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "spec")]
    from tools import p25tx as tx
    rows = []
    for s in range(n):
        st = tx.control_channel(50 + s, 2, lead_idle=10 + 3 * s)
        rows.append(tx.baseband_48k(st.dibits, snr_db=20, seed=s)[0])
    m = min(len(r) for r in rows)
    return np.stack([r[:m] for r in rows])


def _worker(rank, world, port, n_streams, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "spec")]
    import torch.distributed as dist
    from oracle import pyoracle as po
    from p25rx_b200 import shard
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    po.lib().p25o_set_always_correlate(0)
    bb = _streams(n_streams)
    first, count = shard.stream_block(n_streams, rank, world)
    local = [po.MessageReceiver(stream=i).feed(bb[first + i]) for i in range(count)]     # context-local ids
    ev = shard.globalise(np.concatenate(local) if local else np.zeros(0, po.EVENT_DTYPE), first)
    out = shard.gather_events(ev, dst=0)
    if rank == 0:
        q.put(out.view(np.uint8).tobytes())
    dist.barrier()
    dist.destroy_process_group()


def test_stream_block_partition():
    from p25rx_b200 import shard
    for n, w in ((1024, 8), (7, 2), (5, 4), (65536, 8), (3, 8)):
        blocks = [shard.stream_block(n, r, w) for r in range(w)]
        assert sum(c for _, c in blocks) == n
        assert all(blocks[i][0] + blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
        for s in range(0, n, max(1, n // 50)):
            r = shard.owner_of(s, n, w)
            assert blocks[r][0] <= s < blocks[r][0] + blocks[r][1]


@pytest.mark.timeout(180)
def test_two_rank_gather_matches_single_process(oracle):
    n_streams, world = 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_streams, q)) for r in range(world)]
    for p in procs:
        p.start()
    raw = q.get(timeout=150)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got = np.frombuffer(raw, dtype=np.uint8).view(oracle.EVENT_DTYPE)
    bb = _streams(n_streams)
    ref = np.concatenate([oracle.MessageReceiver(stream=s).feed(bb[s]) for s in range(n_streams)])
    from util import events_key
    assert events_key(got) == events_key(ref)
    assert len(got) >= 8 * n_streams - 8
