"""Host adapters of SURVEY.md section 8f rows 1-3 (formats, TSBK/LC fields, stats JSON).  CPU only."""
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "spec"))

import p25_spec as S  # noqa: E402
from p25rx_b200 import consumers as co  # noqa: E402
from p25rx_b200._lib import EVENT_DTYPE  # noqa: E402
from tools import p25tx as tx  # noqa: E402


def test_baseband_file_round_trip_in_replay_blocks():
    x = np.random.default_rng(0).normal(size=20001).astype(np.float32)
    f = io.BytesIO()
    co.write_baseband(f, x)
    f.seek(0)
    blocks = list(co.read_baseband_blocks(f))
    assert [len(b) for b in blocks] == [8192, 8192, 20001 - 16384]        # src/replay.rs:27 block size
    assert np.array_equal(np.concatenate(blocks), x)


def test_iq_chunks_are_the_sdr_buffer_size():
    raw = np.random.default_rng(1).integers(0, 256, 3 * 32768 + 100).astype(np.uint8)
    chunks = list(co.read_iq_chunks(io.BytesIO(raw.tobytes())))
    assert len(chunks) == 3 and all(len(c) == 32768 for c in chunks)       # src/consts.rs:6
    assert np.array_equal(np.concatenate(chunks), raw[: 3 * 32768])


def _ev(kind, payload, stream=0, sample=0):
    e = np.zeros(1, dtype=EVENT_DTYPE)
    e["kind"], e["stream"], e["sample"], e["len"] = kind, stream, sample, len(payload)
    e["payload"][0, : len(payload)] = np.frombuffer(payload, dtype=np.uint8)
    return e


def test_tsbk_fields_match_the_transmitter():
    t = co.TsbkFields(tx.make_tsbk(0x3D, 0, bytes(range(8)), last=True))
    assert t.is_tail and not t.protected and t.opcode() == "ChannelParamsUpdate" and t.mfg() == 0 and t.crc_valid()
    assert t.payload() == bytes(range(8))
    assert co.crc_ccitt_p25(t.raw[:10]) == S.crc_ccitt_p25(t.raw[:10])
    bad = co.TsbkFields(tx.make_tsbk(0x00, 0, bytes(8), last=False, bad_crc=True))
    assert not bad.crc_valid() and co.TsbkFields(tx.make_tsbk(0x11, 0, bytes(8), last=False)).opcode() is None


def test_grants_resolve_to_frequencies_like_recvtask():
    # identifier 1: base 851.00625 MHz, 12.5 kHz spacing, -45 MHz offset, 12.5 kHz bandwidth
    iden = (1 << 60) | (100 << 51) | ((0x000 | 180) << 42) | (100 << 32) | (851_006_250 // 5)
    ev = [_ev(7, tx.make_tsbk(0x00, 0, bytes([0, 0x10, 0x64, 0x12, 0x34, 0, 0, 9]), last=False), 0, 10),     # before IDEN_UP: dropped
          _ev(7, tx.make_tsbk(0x3D, 0, iden.to_bytes(8, "big"), last=False), 0, 20),
          _ev(7, tx.make_tsbk(0x00, 0, bytes([0, 0x10, 0x64, 0x12, 0x34, 0, 0, 9]), last=False), 0, 30),
          _ev(7, tx.make_tsbk(0x02, 0, bytes([0x10, 0x02, 0x00, 0x07, 0x20, 0x03, 0x00, 0x08]), last=True), 0, 40),
          _ev(7, tx.make_tsbk(0x00, 0x90, bytes(8), last=False), 0, 50),                                          # other MFID: ignored
          _ev(7, tx.make_tsbk(0x00, 0, bytes([0, 0x10, 0x01, 0, 5, 0, 0, 0]), last=False, bad_crc=True), 0, 60),  # bad CRC: ignored
          _ev(3, bytes([0x02, 0x10, 0x05, 0x00, 0x09, 0x10, 0x06, 0x00, 0x0A]), 1, 70)]
    chans = co.ChannelParamsMap()
    got = co.collect_talkgroups(np.concatenate(ev), chans)
    p = chans.lookup(1)
    assert p.base_hz == 851_006_250 and p.spacing_hz == 12_500 and p.bandwidth_hz == 12_500 and p.tx_offset_hz == -45_000_000
    assert got == [(0, 30, 0x1234, 851_006_250 + 0x064 * 12_500), (0, 40, 7, 851_006_250 + 2 * 12_500),
                   (1, 70, 9, 851_006_250 + 5 * 12_500), (1, 70, 10, 851_006_250 + 6 * 12_500)]


def test_stats_json_has_the_hub_schema():
    st = np.zeros((12, 4), dtype=np.uint64)
    st[0] = (10, 1, 63, 7)          # bch
    st[10] = (6, 0, 98, 3)          # viterbiDibit
    j = co.stats_json(st)
    assert list(j) == ["bch", "cyclic", "golayStd", "golayExt", "golayShort", "hammingStd", "hammingShort", "rsShort",
                       "rsMed", "rsLong", "viterbiDibit", "viterbiTribit"]                     # src/hub.rs:557-572
    assert j["bch"] == {"totalWords": 10, "errWords": 1, "totalSymbols": 630, "fixedSymbols": 7}   # src/hub.rs:574-581
    assert '"event": "sigPower"' in co.sig_power_event(-31.5)
