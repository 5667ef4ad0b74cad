"""Host-side adapters either side of the hot path.  Integer/byte work on the
host; nothing here computes samples or decodes FEC.

  formats    the reference's on-disk formats: `-w` / `-r` f32le 48 kHz mono baseband (src/main.rs:95-102, :278-283;
             src/replay.rs:26-38) and raw RTL-SDR u8 IQ in BUF_BYTES chunks (src/sdr.rs:25-33, src/consts.rs:6)
  fields     the TSBK / link-control fields the reference's consumers read (src/recv.rs:237-342): opcode, MFID, CRC,
             GroupVoiceGrant, GroupVoiceUpdate (GroupTrafficUpdate), ChannelParamsUpdate and
             ChannelParamsMap::rx_freq (src/recv.rs:330-342).  Layouts are the CAI's [STD], written from memory like
             the rest of spec/ (the p25 crate that holds the reference's parsers is not vendored).
  telemetry  stats and signal power in the hub's JSON schema (src/hub.rs:343, :557-581)
"""
from __future__ import annotations

import json
from dataclasses import dataclass

import numpy as np

from .pipeline import STATS_FAMILIES

BUF_BYTES = 32768           # src/consts.rs:6
BASEBAND_RATE = 48000       # src/consts.rs:13


# --------------------------------------------------------------------------- formats
def write_baseband(f, samples: np.ndarray) -> None:
    """`-w FILE`: f32le / 48 kHz / mono (src/main.rs:99-102, :278-283)."""
    f.write(np.ascontiguousarray(samples, dtype="<f4").tobytes())


def read_baseband_blocks(f, block_bytes: int = BUF_BYTES):
    """`-r FILE` as ReplayReceiver::replay reads it: blocks of up to 32,768 bytes = 8,192 samples
    (src/replay.rs:26-38).  A short final block yields only the samples actually read."""
    while True:
        b = f.read(block_bytes)
        n = len(b) // 4
        if n == 0:
            return
        yield np.frombuffer(b[: 4 * n], dtype="<f4")


def read_iq_chunks(f, chunk_bytes: int = BUF_BYTES):
    """Raw RTL-SDR capture (`rtl_sdr -s 240000`): interleaved u8 I/Q in the 32,768-byte chunks that ReadTask hands
    to DemodTask (src/sdr.rs:25-33); a trailing partial chunk is dropped like a partial USB transfer would be."""
    while True:
        b = f.read(chunk_bytes)
        if len(b) < chunk_bytes:
            return
        yield np.frombuffer(b, dtype=np.uint8)


# --------------------------------------------------------------------------- fields
def crc_ccitt_p25(data: bytes) -> int:
    """TSBK CRC [STD]: x^16 + x^12 + x^5 + 1, zero initial value, result inverted."""
    crc = 0
    for byte in data:
        crc ^= byte << 8
        for _ in range(8):
            crc = ((crc << 1) ^ 0x1021) & 0xFFFF if crc & 0x8000 else (crc << 1) & 0xFFFF
    return crc ^ 0xFFFF


TSBK_OPCODES = {0x00: "GroupVoiceGrant", 0x02: "GroupVoiceUpdate", 0x03: "GroupVoiceUpdateExplicit", 0x04: "UnitVoiceGrant",
                0x2B: "LocRegResponse", 0x2C: "UnitRegResponse", 0x2F: "UnitDeregAck", 0x39: "AltControlChannel",
                0x3A: "RfssStatusBroadcast", 0x3B: "NetworkStatusBroadcast", 0x3C: "AdjacentSite", 0x3D: "ChannelParamsUpdate"}
LC_OPCODES = {0x00: "GroupVoiceTraffic", 0x02: "GroupVoiceUpdate", 0x03: "UnitVoiceTraffic", 0x0F: "CallTermination",
              0x20: "SystemServiceBroadcast", 0x21: "AltControlChannel", 0x22: "AdjacentSite", 0x23: "RfssStatusBroadcast",
              0x24: "NetworkStatusBroadcast"}


@dataclass(frozen=True)
class Channel:
    """4-bit identifier + 12-bit channel number (src/recv.rs:330-337: ch.id(), ch.number())."""
    id: int
    number: int

    @staticmethod
    def from_bits(v: int) -> "Channel":
        return Channel((v >> 12) & 0xF, v & 0xFFF)


@dataclass(frozen=True)
class TsbkFields:
    """The accessors RecvTask::handle_tsbk uses (src/recv.rs:238-266) over the 12 bytes of a TrunkingControl event."""
    raw: bytes

    def __post_init__(self):
        assert len(self.raw) == 12

    @property
    def is_tail(self) -> bool:
        return bool(self.raw[0] & 0x80)

    @property
    def protected(self) -> bool:
        return bool(self.raw[0] & 0x40)

    @property
    def opcode_bits(self) -> int:
        return self.raw[0] & 0x3F

    def opcode(self):
        """Some(opcode) / None like TsbkFields::opcode (src/recv.rs:246-249)."""
        return TSBK_OPCODES.get(self.opcode_bits)

    def mfg(self) -> int:
        return self.raw[1]

    def crc_valid(self) -> bool:
        return crc_ccitt_p25(self.raw[:10]) == (self.raw[10] << 8 | self.raw[11])

    def payload(self) -> bytes:
        return self.raw[2:10]


def group_voice_grant(t: TsbkFields):
    """tsbk::GroupVoiceGrant (src/recv.rs:256-258): service options, channel, talkgroup, source unit."""
    p = t.payload()
    return {"opts": p[0], "channel": Channel.from_bits(p[1] << 8 | p[2]), "talkgroup": p[3] << 8 | p[4],
            "src_unit": p[5] << 16 | p[6] << 8 | p[7]}


def group_traffic_updates(payload: bytes):
    """fields::GroupTrafficUpdate::updates (src/recv.rs:260-262, :300-301, :308-312): two (channel, talkgroup) pairs."""
    return [(Channel.from_bits(payload[0] << 8 | payload[1]), payload[2] << 8 | payload[3]),
            (Channel.from_bits(payload[4] << 8 | payload[5]), payload[6] << 8 | payload[7])]


@dataclass(frozen=True)
class ChannelParams:
    """fields::ChannelParamsUpdate (src/recv.rs:264-266) and ChannelParams::rx_freq (src/recv.rs:336)."""
    id: int
    bandwidth_hz: int
    tx_offset_hz: int
    spacing_hz: int
    base_hz: int

    @staticmethod
    def from_payload(p: bytes) -> "ChannelParams":
        v = int.from_bytes(p, "big")                      # id 4 | bandwidth 9 | offset 9 | spacing 10 | base 32
        ident = v >> 60
        bw = (v >> 51) & 0x1FF
        off = (v >> 42) & 0x1FF
        spacing = (v >> 32) & 0x3FF
        base = v & 0xFFFFFFFF
        sign = 1 if off & 0x100 else -1
        return ChannelParams(ident, bw * 125, sign * (off & 0xFF) * 250_000, spacing * 125, base * 5)

    def rx_freq(self, number: int) -> int:
        return self.base_hz + self.spacing_hz * number


class ChannelParamsMap:
    """`self.channels` of RecvTask (src/recv.rs:265, :335-338): identifier -> parameters."""

    def __init__(self):
        self._m: dict[int, ChannelParams] = {}

    def update(self, p: ChannelParams) -> None:
        self._m[p.id] = p

    def lookup(self, ident: int):
        return self._m.get(ident)


@dataclass(frozen=True)
class LinkControlFields:
    """LinkControlFields::{opcode, payload} (src/recv.rs:280, :298) over the 9 bytes of LinkControl / VoiceTerm."""
    raw: bytes

    def __post_init__(self):
        assert len(self.raw) == 9

    def opcode(self):
        return LC_OPCODES.get(self.raw[0] & 0x3F)

    def payload(self) -> bytes:
        return self.raw[1:9]


def _be(p: bytes) -> int:
    return int.from_bytes(p, "big")


def hub_status_event(kind: int, raw: bytes, channels: "ChannelParamsMap"):
    """The JSON objects the hub streams for one TrunkingControl (kind 7) or LinkControl / VoiceTerm (3 / 8) event
    (src/hub.rs:346-404, :405-443, :525-547): a list of (event name, payload).  Same [STD]-from-memory layouts as
    include/p25cu.hpp `fields::`."""
    out = []

    def freq(ch: Channel):
        p = channels.lookup(ch.id)
        return None if p is None else p.rx_freq(ch.number)

    def rfss(p):
        out.append(("rfssStatus", {"area": p[0], "system": (p[1] & 0xF) << 8 | p[2], "rfss": p[3], "site": p[4]}))

    def net(p):
        out.append(("networkStatus", {"area": p[0], "wacn": p[1] << 12 | p[2] << 4 | p[3] >> 4, "system": (p[3] & 0xF) << 8 | p[4]}))

    def adjacent(p):
        f = freq(Channel.from_bits(_be(p[5:7])))
        if f is not None:
            out.append(("adjacentSite", {"area": p[0], "rfss": p[3], "system": (p[1] & 0xF) << 8 | p[2], "site": p[4], "freq": f}))

    def alt(p):
        for ch in (Channel.from_bits(_be(p[2:4])), Channel.from_bits(_be(p[5:7]))):
            f = freq(ch)
            if f is not None:
                out.append(("altControl", {"rfss": p[0], "site": p[1], "freq": f}))

    if kind == 7:
        t = TsbkFields(raw[:12])
        if t.mfg() != 0 or not t.crc_valid() or t.opcode() is None:
            return out
        p, op = t.payload(), t.opcode()
        if op == "RfssStatusBroadcast":
            rfss(p)
        elif op == "NetworkStatusBroadcast":
            net(p)
        elif op == "AltControlChannel":
            alt(p)
        elif op == "AdjacentSite":
            adjacent(p)
        elif op == "LocRegResponse":
            out.append(("locReg", {"response": p[0] & 3, "rfss": p[3], "site": p[4], "unit": _be(p[5:8])}))
        elif op == "UnitRegResponse":
            out.append(("unitReg", {"response": (p[0] >> 4) & 3, "system": (p[0] & 0xF) << 8 | p[1], "unitId": _be(p[2:5]), "unitAddr": _be(p[5:8])}))
        elif op == "UnitDeregAck":
            out.append(("unitDereg", {"wacn": p[1] << 12 | p[2] << 4 | p[3] >> 4, "system": (p[3] & 0xF) << 8 | p[4], "unit": _be(p[5:8])}))
    elif kind in (3, 8):
        lc = LinkControlFields(raw[:9])
        op, p = lc.opcode(), lc.payload()
        if op == "GroupVoiceTraffic":
            out.append(("srcUnit", _be(raw[6:9])))
        elif op == "RfssStatusBroadcast":
            rfss(p)
        elif op == "NetworkStatusBroadcast":
            net(p)
        elif op == "AdjacentSite":
            adjacent(p)
        elif op == "AltControlChannel":
            alt(p)
    return out


def talkgroup_other(tg: int) -> bool:
    """TalkGroup::Other(_) (src/recv.rs:327-330): not Nobody (0x0000), Default (0x0001) or Everybody (0xFFFF)."""
    return tg not in (0x0000, 0x0001, 0xFFFF)


def consume(events: np.ndarray):
    """One RecvTask + hub per stream over a drained event array: returns (talkgroups [(stream, sample, tg, rx_freq)],
    hub [(stream, json text)]) in the order tests/cpp/p25host_main prints them in its `consumer` mode."""
    tgs, hub = [], []
    for s in sorted(set(int(x) for x in events["stream"])):
        ch = ChannelParamsMap()
        ev = events[events["stream"] == s]
        tgs += collect_talkgroups(ev, ch, hub_out=hub)
    return tgs, hub


def collect_talkgroups(events: np.ndarray, channels: ChannelParamsMap, hub_out: list | None = None):
    """What RecvTask::handle_tsbk / handle_lc / add_talkgroup make of a drained event array (src/recv.rs:237-342):
    returns [(stream, sample, talkgroup, rx_freq_hz)] for every grant / update whose channel identifier is known."""
    out = []

    def add(ev, tg, ch):
        p = channels.lookup(ch.id)
        if p is not None and talkgroup_other(tg):
            out.append((int(ev["stream"]), int(ev["sample"]), tg, p.rx_freq(ch.number)))

    for ev in events:
        kind = int(ev["kind"])
        if hub_out is not None and kind in (3, 7, 8):
            pending_hub = [(int(ev["stream"]), json.dumps({"event": n, "payload": pl})) for n, pl in
                           hub_status_event(kind, bytes(ev["payload"][:12]), channels)]
        else:
            pending_hub = []
        if kind == 7:
            t = TsbkFields(bytes(ev["payload"][:12]))
            if t.mfg() != 0 or not t.crc_valid():
                continue
            op = t.opcode()
            if op == "ChannelParamsUpdate":
                channels.update(ChannelParams.from_payload(t.payload()))
            elif op == "GroupVoiceGrant":
                g = group_voice_grant(t)
                add(ev, g["talkgroup"], g["channel"])
            elif op == "GroupVoiceUpdate":
                for ch, tg in group_traffic_updates(t.payload()):
                    add(ev, tg, ch)
        elif kind in (3, 8):      # LinkControl and VoiceTerm(lc) both go through handle_lc (src/recv.rs:229, :232)
            lc = LinkControlFields(bytes(ev["payload"][:9]))
            if lc.opcode() == "GroupVoiceUpdate":
                for ch, tg in group_traffic_updates(lc.payload()):
                    add(ev, tg, ch)
        if hub_out is not None:
            hub_out += pending_hub
    return out


# --------------------------------------------------------------------------- telemetry
def stats_json(stats: np.ndarray) -> dict:
    """serialize_stats / serialize_code_stats (src/hub.rs:557-581) for one stream's 12 x {words, errs, size, fixed}."""
    return {name: {"totalWords": int(w), "errWords": int(e), "totalSymbols": int(w) * int(sz), "fixedSymbols": int(fx)}
            for name, (w, e, sz, fx) in zip(STATS_FAMILIES, np.asarray(stats).reshape(12, 4))}


def sse_event(name: str, payload) -> str:
    """One server-sent event in the hub's framing: {"event": name, "payload": ...} (src/hub.rs SerdeEvent)."""
    return "data: " + json.dumps({"event": name, "payload": payload}) + "\n\n"


def sig_power_event(power_dbm: float) -> str:
    """HubEvent::UpdateSignalPower -> "sigPower" (src/demod.rs:95-101, src/hub.rs:343)."""
    return sse_event("sigPower", float(power_dbm))
