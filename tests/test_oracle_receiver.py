"""Oracle MessageReceiver against the synthetic transmitter's ground truth (CPU only).

The reference has no hot-path tests; these follow its replay shape
(src/replay.rs:40-57): a stream of f32 baseband samples in, MessageEvents out."""
import numpy as np
import pytest
from tools import p25tx as tx
from util import check_against_truth, events_key


def test_control_channel_clean(oracle):
    st = tx.control_channel(1, 5)
    bb, _ = tx.baseband_48k(st.dibits)
    rx = oracle.MessageReceiver()
    ev = rx.feed(bb)
    check_against_truth(ev, tx.expected_events(st))
    stats = rx.stats()
    assert stats[0, 0] == 5 and stats[10, 0] == 15 and stats[:, 1].sum() == 0


@pytest.mark.parametrize("snr,dc,toff", [(25, 0.0, 0.0), (15, 0.02, 3.3), (12, -0.03, 7.7)])
def test_control_channel_noise_dc_timing(oracle, snr, dc, toff):
    st = tx.control_channel(2, 6)
    bb, _ = tx.baseband_48k(st.dibits, snr_db=snr, dc=dc, seed=9, timing_offset=toff)
    check_against_truth(oracle.MessageReceiver().feed(bb), tx.expected_events(st))


def test_traffic_channel(oracle):
    st = tx.traffic_channel(3, 2)
    bb, _ = tx.baseband_48k(st.dibits, snr_db=26, seed=4, dc=0.03)
    rx = oracle.MessageReceiver()
    ev = rx.feed(bb)
    check_against_truth(ev, tx.expected_events(st))
    kinds = set(ev["kind"].tolist())
    assert {tx.EV_NID, tx.EV_VOICE_HEADER, tx.EV_LINK_CONTROL, tx.EV_CRYPTO_CONTROL, tx.EV_LSD, tx.EV_VOICE_FRAME,
            tx.EV_VOICE_TERM} <= kinds
    s = rx.stats()
    assert s[2, 0] == 4 * 36 and s[5, 0] == 3 * 36 and s[6, 0] == 4 * 24 and s[1, 0] == 8
    assert s[4, 0] == 36 and s[3, 0] == 12 and s[9, 0] == 1 and s[7, 0] == 3 and s[8, 0] == 2


def test_chunking_invariance(oracle):
    """feed() is sample-at-a-time: any chunking gives the same events at the same sample indices."""
    st = tx.traffic_channel(5, 1)
    bb, _ = tx.baseband_48k(st.dibits, snr_db=20, seed=2)
    whole = oracle.MessageReceiver().feed(bb)
    for chunk in (1, 7, 3277, 8192):
        rx = oracle.MessageReceiver()
        parts = [rx.feed(bb[i:i + chunk]) for i in range(0, len(bb), chunk)]
        assert events_key(np.concatenate(parts)) == events_key(whole)


def test_event_sample_index_is_the_completing_symbol(oracle):
    st = tx.control_channel(7, 2, lead_idle=60)
    bb, centres = tx.baseband_48k(st.dibits)
    ev = oracle.MessageReceiver().feed(bb)
    truth = sorted(st.events, key=lambda e: e[2])
    for e, (_, _, dibit_idx) in zip(ev, truth):
        assert abs(float(e["sample"]) - centres[dibit_idx]) <= 1.0


def test_error_paths(oracle):
    # 1. NID destroyed -> BCH error, receiver resyncs and decodes the next unit
    st = tx.control_channel(11, 3)
    d = st.dibits.copy()
    d[40 + 24:40 + 24 + 33] = np.random.default_rng(1).integers(0, 4, 33)
    bb, _ = tx.baseband_48k(d)
    ev = oracle.MessageReceiver().feed(bb)
    assert int(ev[0]["kind"]) == tx.EV_ERROR and int(ev[0]["payload"][0]) == 1
    assert [int(k) for k in ev["kind"][1:5]] == [tx.EV_NID, tx.EV_TSBK, tx.EV_TSBK, tx.EV_TSBK]
    # 2. a TSBK block destroyed -> Viterbi error
    d = st.dibits.copy()
    d[40 + 60:40 + 60 + 90] = np.random.default_rng(0).integers(0, 4, 90)
    bb, _ = tx.baseband_48k(d)
    ev = oracle.MessageReceiver().feed(bb)
    assert [int(k) for k in ev["kind"][:2]] == [tx.EV_NID, tx.EV_ERROR] and int(ev[1]["payload"][0]) == 3
    # 3. CRC failures are still delivered (the consumer checks, reference src/recv.rs:242)
    u = tx.tsdu(0x293, [tx.make_tsbk(0x3A, 0, bytes(8), last=True, bad_crc=True)])
    bb, _ = tx.baseband_48k(tx.concat_units([u], lead_idle=30).dibits)
    ev = oracle.MessageReceiver().feed(bb)
    assert int(ev[1]["kind"]) == tx.EV_TSBK
    # 4. packet data units and simple terminators produce a NID only
    rng = np.random.default_rng(4)
    pdu = tx.pdu(0x293, True, [tx.confirmed_block(j, rng.integers(0, 256, 16).astype(np.uint8).tobytes()) for j in range(3)])
    for duid, un in ((0xC, pdu), (0x3, tx._assemble(0x293, 0x3, np.zeros(0, np.uint8), []))):
        bb, _ = tx.baseband_48k(tx.concat_units([un, tx.tsdu(0x293, [tx.make_tsbk(1, 0, bytes(8), True)])], lead_idle=30, gap_idle=300).dibits)
        ev = oracle.MessageReceiver().feed(bb)
        assert [int(k) for k in ev["kind"]] == [tx.EV_NID, tx.EV_NID, tx.EV_TSBK]
        assert int(ev[0]["payload"][2]) == duid
    # 5. reserved DUID -> UnknownNID error
    un = tx._assemble(0x293, 0x9, np.zeros(0, np.uint8), [])
    bb, _ = tx.baseband_48k(tx.concat_units([un], lead_idle=30).dibits)
    ev = oracle.MessageReceiver().feed(bb)
    assert int(ev[0]["kind"]) == tx.EV_ERROR and int(ev[0]["payload"][0]) == 5


def test_resync_drops_lock(oracle):
    st = tx.control_channel(13, 2)
    bb, _ = tx.baseband_48k(st.dibits)
    rx = oracle.MessageReceiver()
    cut = 40 * 10 + 80 + 600        # inside the first NID/TSBK
    ev1 = rx.feed(bb[:cut])
    assert rx.state != 0
    rx.resync()
    assert rx.state == 0
    ev2 = rx.feed(bb[cut:])
    # the interrupted TSDU is lost, the second one decodes in full
    assert [int(k) for k in ev2["kind"]] == [tx.EV_NID, tx.EV_TSBK, tx.EV_TSBK, tx.EV_TSBK]
    assert len(ev1) <= 1


def test_noise_only_produces_no_lock(oracle):
    rng = np.random.default_rng(5)
    bb = rng.normal(0, 0.5, 48000).astype(np.float32)
    ev = oracle.MessageReceiver().feed(bb)
    assert len(ev) <= 2      # false locks on noise are possible but rare; they can only yield errors
    assert all(int(k) == tx.EV_ERROR for k in ev["kind"])


def test_packet_data_units(oracle):
    """a PDU yields its PacketNID only (no MessageEvent variant carries packet data, src/recv.rs:214-233);
    header and unconfirmed blocks count as viterbiDibit words, confirmed blocks as viterbiTribit words (src/hub.rs:569-570);
    a destroyed confirmed block is a ViterbiTribit error; a header whose CRC fails drops the lock silently."""
    st = tx.data_channel(21, 6)
    bb, _ = tx.baseband_48k(st.dibits, snr_db=20, seed=1)
    rx = oracle.MessageReceiver()
    ev = rx.feed(bb)
    check_against_truth(ev, tx.expected_events(st))
    stats = rx.stats()
    n_conf = sum(1 for i in range(0, 6, 2))          # PDUs 0, 2, 4 are confirmed
    assert stats[11, 0] > 0 and stats[11, 1] == 0 and stats[10, 0] > 6      # tribit words seen, none failed
    # destroy one confirmed data block: Error(ViterbiTribit), then resync and the following TSDU decodes
    rng = np.random.default_rng(2)
    pdu = tx.pdu(0x293, True, [tx.confirmed_block(j, bytes(16)) for j in range(2)])
    d = tx.concat_units([pdu, tx.tsdu(0x293, [tx.make_tsbk(3, 0, bytes(8), True)])], lead_idle=30, gap_idle=40).dibits.copy()
    blk = 30 + 24 + 32 + 98 + 98 + 6             # inside the second data block (status symbols shift it a little)
    d[blk + 10: blk + 80] = rng.integers(0, 4, 70)
    ev = oracle.MessageReceiver().feed(tx.baseband_48k(d)[0])
    assert [int(k) for k in ev["kind"]] == [tx.EV_NID, tx.EV_ERROR, tx.EV_NID, tx.EV_TSBK] and int(ev[1]["payload"][0]) == 4
    # header CRC failure: no event besides the NID, lock dropped, next unit decodes
    bad = tx.pdu(0x293, True, [tx.confirmed_block(0, bytes(16))], bad_header_crc=True)
    d = tx.concat_units([bad, tx.tsdu(0x293, [tx.make_tsbk(3, 0, bytes(8), True)])], lead_idle=30, gap_idle=200).dibits
    ev = oracle.MessageReceiver().feed(tx.baseband_48k(d)[0])
    assert [int(k) for k in ev["kind"]] == [tx.EV_NID, tx.EV_NID, tx.EV_TSBK]
