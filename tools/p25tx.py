"""Synthetic P25 Phase 1 transmitter (numpy) -- test and bench input generator.

Builds data units (TSDU, HDU, LDU1, LDU2, TDU, TDULC) from payload fields with the
encoders of spec/p25_spec.py, inserts status symbols, shapes the dibits as C4FM and
produces either ideal 48 kHz baseband (the `-r FILE` replay format of the reference,
src/main.rs:95-98, src/replay.rs:26-38) or FM-modulated IQ (cf32 or RTL-SDR style u8,
src/sdr.rs:25-33) with AWGN, carrier offset and timing offset.  Every generator returns
the ground-truth event list next to the samples.

The reference ships no transmitter and no captures; this file is new.
"""
from __future__ import annotations

import os
import sys
from dataclasses import dataclass, field

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "spec"))
import p25_spec as S  # noqa: E402

EV_ERROR, EV_NID, EV_VOICE_HEADER, EV_LINK_CONTROL, EV_CRYPTO_CONTROL, EV_LSD, EV_VOICE_FRAME, EV_TSBK, EV_VOICE_TERM = range(9)


# --------------------------------------------------------------------------- bits
def _bits_msb(value: int, n: int) -> list[int]:
    return [(value >> (n - 1 - i)) & 1 for i in range(n)]


def _bits_to_dibits(bits) -> np.ndarray:
    bits = np.asarray(bits, dtype=np.uint8)
    assert len(bits) % 2 == 0
    return (bits[0::2] << 1) | bits[1::2]


def _bytes_to_hexbits(data: bytes) -> list[int]:
    bits = np.unpackbits(np.frombuffer(bytes(data), dtype=np.uint8))
    assert len(bits) % 6 == 0
    return [int("".join(map(str, bits[i:i + 6])), 2) for i in range(0, len(bits), 6)]


# --------------------------------------------------------------------------- data units
@dataclass
class Unit:
    """One data unit: transmitted dibits (status symbols included) + expected events.

    Each expected event is (kind, payload bytes, dibit index within the unit at which the
    receiver completes it)."""
    dibits: np.ndarray
    events: list = field(default_factory=list)


def _assemble(nac: int, duid: int, payload: np.ndarray, marks: list, status: int = 0b10) -> Unit:
    """FS + NID + payload + null pad to a status boundary, status symbol every 36th dibit."""
    nid = S.nid_encode(nac, duid)
    data = np.concatenate([S.frame_sync_dibits(), _bits_to_dibits(_bits_msb(nid, 64)), payload.astype(np.uint8)])
    pad = (-len(data)) % 35
    data = np.concatenate([data, np.zeros(pad, dtype=np.uint8)])
    nblk = len(data) // 35
    out = np.empty(nblk * 36, dtype=np.uint8)
    out.reshape(nblk, 36)[:, :35] = data.reshape(nblk, 35)
    out.reshape(nblk, 36)[:, 35] = status

    def tx_index(data_idx: int) -> int:  # index in `out` of data dibit `data_idx`
        return data_idx + data_idx // 35

    nac_b = bytes([nac & 0xFF, (nac >> 8) & 0xF, duid])
    events = [(EV_NID, nac_b, tx_index(24 + 32 - 1))]
    for kind, pl, payload_dibit_end in marks:
        events.append((kind, pl, tx_index(56 + payload_dibit_end - 1)))
    return Unit(out, events)


def make_tsbk(opcode: int, mfid: int, args: bytes, last: bool, protect: bool = False, bad_crc: bool = False) -> bytes:
    assert len(args) == 8
    head = bytes([(0x80 if last else 0) | (0x40 if protect else 0) | (opcode & 0x3F), mfid & 0xFF]) + bytes(args)
    crc = S.crc_ccitt_p25(head)
    if bad_crc:
        crc ^= 0x0101
    return head + bytes([crc >> 8, crc & 0xFF])


def tsdu(nac: int, tsbks: list[bytes]) -> Unit:
    assert 1 <= len(tsbks) <= 3
    payload = np.concatenate([S.tsbk_block_dibits(t) for t in tsbks])
    marks = [(EV_TSBK, bytes(t), 98 * (i + 1)) for i, t in enumerate(tsbks)]
    return _assemble(nac, S.DUID_TSDU, payload, marks)


def hdu(nac: int, mi: bytes, mfid: int, algid: int, kid: int, tgid: int) -> Unit:
    assert len(mi) == 9
    fields = bytes(mi) + bytes([mfid, algid, kid >> 8, kid & 0xFF, tgid >> 8, tgid & 0xFF])
    hexbits = S.rs_encode(_bytes_to_hexbits(fields), *S.RS_LONG)
    bits = []
    for h in hexbits:
        bits += _bits_msb(S.golay18_encode(h), 18)
    return _assemble(nac, S.DUID_HDU, _bits_to_dibits(bits), [(EV_VOICE_HEADER, fields, S.HDU_DIBITS)])


def tdu(nac: int) -> Unit:
    return _assemble(nac, S.DUID_TDU, np.zeros(0, dtype=np.uint8), [])


def tdulc(nac: int, lc: bytes) -> Unit:
    assert len(lc) == 9
    hexbits = S.rs_encode(_bytes_to_hexbits(lc), *S.RS_SHORT)
    bits = []
    for i in range(12):
        bits += _bits_msb(S.golay24_encode((hexbits[2 * i] << 6) | hexbits[2 * i + 1]), 24)
    return _assemble(nac, S.DUID_TDULC, _bits_to_dibits(bits), [(EV_VOICE_TERM, bytes(lc), S.TDULC_DIBITS)])


def pdu(nac: int, confirmed: bool, blocks: list[bytes], bad_header_crc: bool = False, announce: int | None = None) -> Unit:
    """Packet data unit [STD]: 1/2-rate header block, then `blocks` -- 18-byte confirmed-data blocks (3/4-rate) or
    12-byte unconfirmed ones (1/2-rate).  The receiver surface has no event for packet data (src/recv.rs:214-233): the
    only expected event is the PacketNID.  announce overrides the header's blocks-to-follow field."""
    n_follow = len(blocks) if announce is None else announce
    head = S.pdu_header(S.PDU_FORMAT_CONFIRMED if confirmed else S.PDU_FORMAT_UNCONFIRMED, n_follow, bad_crc=bad_header_crc)
    parts = [S.tsbk_block_dibits(head)]
    for b in blocks:
        parts.append(S.pdu_block34_dibits(b) if confirmed else S.tsbk_block_dibits(b))
    return _assemble(nac, S.DUID_PDU, np.concatenate(parts), [])


def confirmed_block(serial: int, data16: bytes) -> bytes:
    """18-byte confirmed data block: 7-bit serial number, CRC-9 over serial + data, 16 data octets [STD layout, RECALL]."""
    assert len(data16) == 16
    bits = _bits_msb(serial & 0x7F, 7) + list(np.unpackbits(np.frombuffer(bytes(data16), dtype=np.uint8)))
    crc = S.crc9_p25(bits)
    return bytes([((serial & 0x7F) << 1) | (crc >> 8), crc & 0xFF]) + bytes(data16)


def data_channel(seed: int, n_pdu: int, nac: int = 0x293, lead_idle: int = 40) -> Stream:
    """Packet data traffic: confirmed and unconfirmed PDUs of 1..4 blocks between TSDUs (exercises the 3/4-rate trellis,
    SURVEY.md 8a a9.9)."""
    rng = np.random.default_rng(seed)
    units = []
    for i in range(n_pdu):
        nb = int(rng.integers(1, 5))
        if i % 2 == 0:
            blocks = [confirmed_block(j, rng.integers(0, 256, 16).astype(np.uint8).tobytes()) for j in range(nb)]
            units.append(pdu(nac, True, blocks))
        else:
            units.append(pdu(nac, False, [rng.integers(0, 256, 12).astype(np.uint8).tobytes() for _ in range(nb)]))
        units.append(tsdu(nac, [make_tsbk(int(rng.integers(0, 64)), 0, rng.integers(0, 256, 8).astype(np.uint8).tobytes(), last=True)]))
    return concat_units(units, lead_idle=lead_idle, gap_idle=6, rng=rng)


def _vf_payload(u: list[int]) -> bytes:
    return np.array(list(u) + [0] * 7, dtype="<u4").tobytes()


def ldu(nac: int, which: int, frames: list[list[int]], word: bytes, lsd: tuple[int, int]) -> Unit:
    """which = 1: `word` is the 9-byte link control; which = 2: the 12-byte crypto sync."""
    assert len(frames) == 9
    if which == 1:
        assert len(word) == 9
        hexbits = S.rs_encode(_bytes_to_hexbits(word), *S.RS_SHORT)
        kind, duid = EV_LINK_CONTROL, S.DUID_LDU1
    else:
        assert len(word) == 12
        hexbits = S.rs_encode(_bytes_to_hexbits(word), *S.RS_MED)
        kind, duid = EV_CRYPTO_CONTROL, S.DUID_LDU2
    parts, marks = [], []
    vf = chunk = 0
    off = 0
    for pkind, n in S.LDU_LAYOUT:
        if pkind == "vf":
            parts.append(S.imbe_encode(frames[vf]))
            marks.append((EV_VOICE_FRAME, _vf_payload(frames[vf]), off + n))
            vf += 1
        elif pkind == "lc":
            bits = []
            for h in hexbits[4 * chunk:4 * chunk + 4]:
                bits += _bits_msb(S.hamming10_encode(h), 10)
            parts.append(_bits_to_dibits(bits))
            chunk += 1
            if chunk == 6:
                marks.append((kind, bytes(word), off + n))
        else:
            bits = _bits_msb(S.cyclic16_encode(lsd[0]), 16) + _bits_msb(S.cyclic16_encode(lsd[1]), 16)
            parts.append(_bits_to_dibits(bits))
            marks.append((EV_LSD, np.array([(lsd[0] << 8) | lsd[1]], dtype="<u4").tobytes(), off + n))
        off += n
    return _assemble(nac, duid, np.concatenate(parts), marks)


def random_imbe(rng: np.random.Generator) -> list[int]:
    return [int(rng.integers(0, 1 << b)) for b in S.IMBE_U_BITS]


# --------------------------------------------------------------------------- streams
@dataclass
class Stream:
    dibits: np.ndarray
    events: list          # (kind, payload, dibit index in `dibits`)


def concat_units(units: list[Unit], lead_idle: int = 0, gap_idle: int = 0, rng=None) -> Stream:
    """Join units; idle gaps are random dibits (an unsynchronised carrier)."""
    rng = rng or np.random.default_rng(0)
    parts, events, off = [], [], 0

    def idle(n):
        nonlocal off
        if n:
            parts.append(rng.integers(0, 4, n).astype(np.uint8))
            off += n

    idle(lead_idle)
    for i, u in enumerate(units):
        events += [(k, p, off + e) for k, p, e in u.events]
        parts.append(u.dibits)
        off += len(u.dibits)
        if i + 1 < len(units):
            idle(gap_idle)
    return Stream(np.concatenate(parts), events)


def control_channel(seed: int, n_tsdu: int, nac: int = 0x293, lead_idle: int = 40) -> Stream:
    """Back-to-back triple-TSBK TSDUs with random opcodes/arguments and valid CRCs
    (SURVEY.md section 8d, cfg1/cfg2)."""
    rng = np.random.default_rng(seed)
    units = []
    for _ in range(n_tsdu):
        blocks = [make_tsbk(int(rng.integers(0, 64)), 0, rng.integers(0, 256, 8).astype(np.uint8).tobytes(), last=(b == 2))
                  for b in range(3)]
        units.append(tsdu(nac, blocks))
    return concat_units(units, lead_idle=lead_idle, rng=rng)


def traffic_channel(seed: int, n_ldu_pairs: int, nac: int = 0x293, lead_idle: int = 40) -> Stream:
    """HDU, alternating LDU1/LDU2, TDULC."""
    rng = np.random.default_rng(seed)
    tgid = int(rng.integers(1, 65535))
    units = [hdu(nac, bytes(9), 0, 0x80, 0, tgid)]
    for _ in range(n_ldu_pairs):
        lc = bytes([0x00, 0x00, 0x00, 0x00, tgid >> 8, tgid & 0xFF]) + rng.integers(0, 256, 3).astype(np.uint8).tobytes()
        units.append(ldu(nac, 1, [random_imbe(rng) for _ in range(9)], lc,
                         (int(rng.integers(0, 256)), int(rng.integers(0, 256)))))
        es = bytes(9) + bytes([0x80, 0x00, 0x00])
        units.append(ldu(nac, 2, [random_imbe(rng) for _ in range(9)], es,
                         (int(rng.integers(0, 256)), int(rng.integers(0, 256)))))
    units.append(tdulc(nac, bytes([0x0F, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x01])))
    units.append(tdu(nac))
    return concat_units(units, lead_idle=lead_idle, rng=rng)


# --------------------------------------------------------------------------- modulation
_SYM = np.array([+1.0, +3.0, -1.0, -3.0])      # indexed by dibit 00, 01, 10, 11  [STD]


def dibits_to_symbols(dibits: np.ndarray) -> np.ndarray:
    return _SYM[np.asarray(dibits, dtype=np.int64)]


def _c4fm_pulse_48k(span: int = 8) -> np.ndarray:
    """Impulse response of H(f) P(f) [STD C4FM: raised cosine x inverse sinc] at 48 kHz."""
    n = np.arange(-span * S.SPS, span * S.SPS + 1, dtype=np.float64)
    f = np.linspace(0.0, 2880.0, 2881)
    H = np.where(f < 1920.0, 1.0, 0.5 + 0.5 * np.cos(2 * np.pi * f / 1920.0))
    x = np.pi * f / 4800.0
    P = np.where(f == 0, 1.0, x / np.sin(np.where(f == 0, 1.0, x)))
    resp = H * P
    h = np.array([2.0 * np.trapezoid(resp * np.cos(2 * np.pi * f * k / 48000.0), f) for k in n]) / 48000.0
    h *= np.hanning(len(h) + 2)[1:-1] ** 0.25
    return h * S.SPS      # unit symbol impulse -> unit-area response at 10 samples/symbol


_PULSE = None


def deviation_48k(dibits: np.ndarray, timing_offset: float = 0.0) -> np.ndarray:
    """Instantaneous frequency (Hz) of the C4FM signal at 48 kHz, 10 samples per symbol.

    Symbol i is centred on sample 10*i + 80 + timing_offset (80 = filter delay)."""
    global _PULSE
    if _PULSE is None:
        _PULSE = _c4fm_pulse_48k()
    sym = dibits_to_symbols(dibits) * S.SYMBOL_DEVIATION_HZ
    up = np.zeros(len(sym) * S.SPS)
    up[:: S.SPS] = sym
    from scipy import signal
    dev = signal.fftconvolve(up, _PULSE)
    if timing_offset:
        k = np.arange(len(dev)) - timing_offset
        dev = np.interp(k, np.arange(len(dev)), dev, left=0.0, right=0.0)
    return dev


def baseband_48k(dibits: np.ndarray, snr_db: float | None = None, dc: float = 0.0, seed: int = 0,
                 timing_offset: float = 0.0) -> tuple[np.ndarray, np.ndarray]:
    """Ideal discriminator + integrator output (f32, the reference's replay format).

    Returns (samples, symbol_centres): symbol_centres[i] is the (fractional) sample index of
    the centre of dibit i after the 10-sample integrator (delay 4.5 samples)."""
    dev = deviation_48k(dibits, timing_offset)
    x = dev / S.FM_DEVIATION_HZ                       # discriminator gain fs/(2 pi 5000) on 2 pi f/fs
    box = np.convolve(x, np.ones(S.BOXCAR_LEN) / S.BOXCAR_LEN)[: len(x)]
    if snr_db is not None:
        rng = np.random.default_rng(seed)
        p = 5.0 * (S.SYMBOL_DEVIATION_HZ / S.FM_DEVIATION_HZ) ** 2   # mean symbol power of +-1,+-3
        box = box + rng.normal(0.0, np.sqrt(p / 10 ** (snr_db / 10)), len(box))
    centres = np.arange(len(dibits)) * S.SPS + 8 * S.SPS + 4.5 + timing_offset
    return (box + dc).astype(np.float32), centres


def modulate_iq(dibits: np.ndarray, fs: int, snr_db: float | None = 30.0, cfo_hz: float = 0.0,
                timing_offset: float = 0.0, seed: int = 0, amplitude: float = 0.5) -> np.ndarray:
    """FM-modulated complex baseband at `fs` (240 kS/s or 2.4 MS/s), complex64.

    snr_db is carrier power over noise power in a 12.5 kHz channel."""
    from scipy import signal
    assert fs % S.BASEBAND_SAMPLE_RATE == 0
    L = fs // S.BASEBAND_SAMPLE_RATE
    dev = deviation_48k(dibits, timing_offset)
    if L > 1:
        dev = signal.resample_poly(dev, L, 1, window=("kaiser", 8.0))
    phase = 2.0 * np.pi * np.cumsum(dev + cfo_hz) / fs
    rng = np.random.default_rng(seed)
    iq = amplitude * np.exp(1j * (phase + rng.uniform(0, 2 * np.pi)))
    if snr_db is not None:
        sigma2 = amplitude ** 2 / 10 ** (snr_db / 10) * (fs / 12500.0)
        iq = iq + np.sqrt(sigma2 / 2) * (rng.standard_normal(len(iq)) + 1j * rng.standard_normal(len(iq)))
    return iq.astype(np.complex64)


def modulate_iq_periodic(dibits: np.ndarray, fs: int, snr_db: float | None = 30.0, cfo_cycles: int = 0, seed: int = 0,
                          amplitude: float = 0.5) -> np.ndarray:
    """One period of a seamlessly repeating C4FM signal: the symbol shaping is circular and the
    carrier offset is trimmed so that the phase advances by a whole number of turns per period.
    Feeding the returned buffer over and over is a continuous transmission of `dibits` repeated
    (bench.py uses this so that every timed step sees a valid, phase-continuous signal).
    cfo_cycles: additional whole carrier turns per period (CFO = cfo_cycles / period)."""
    global _PULSE
    if _PULSE is None:
        _PULSE = _c4fm_pulse_48k()
    assert fs % S.BASEBAND_SAMPLE_RATE == 0
    L = fs // S.BASEBAND_SAMPLE_RATE
    n48 = len(dibits) * S.SPS
    up = np.zeros(n48)
    up[:: S.SPS] = dibits_to_symbols(dibits) * S.SYMBOL_DEVIATION_HZ
    pulse = np.zeros(n48)
    half = len(_PULSE) // 2
    pulse[: half + 1] = _PULSE[half:]
    pulse[-half:] = _PULSE[:half]
    spec = np.fft.rfft(up) * np.fft.rfft(pulse)                  # circular shaping at 48 kHz
    full = np.zeros(n48 * L // 2 + 1, dtype=np.complex128)
    full[: len(spec)] = spec
    if n48 % 2 == 0:
        full[len(spec) - 1] *= 0.5
    dev = np.fft.irfft(full, n48 * L) * L                         # band-limited interpolation to fs
    turns = np.sum(dev) / fs                                      # carrier turns per period from the data
    trim = (np.round(turns) - turns + cfo_cycles) * fs / len(dev)  # Hz
    phase = 2.0 * np.pi * np.cumsum(dev + trim) / fs
    rng = np.random.default_rng(seed)
    iq = amplitude * np.exp(1j * (phase + rng.uniform(0, 2 * np.pi)))
    if snr_db is not None:
        sigma2 = amplitude ** 2 / 10 ** (snr_db / 10) * (fs / 12500.0)
        iq = iq + np.sqrt(sigma2 / 2) * (rng.standard_normal(len(iq)) + 1j * rng.standard_normal(len(iq)))
    return iq.astype(np.complex64)


def wideband_capture(channels: dict, n_samples: int, noise_db: float | None = -50.0, seed: int = 0) -> np.ndarray:
    """A 19.2 MS/s cf32 capture: channels maps a channel number k (0..1535, centre
    k * 12.5 kHz above the capture centre, k >= 768 below it) to (dibits, amplitude, extra_cfo_hz); white noise of
    noise_db dBFS total power is added over the whole band."""
    fs = S.PFB_SAMPLE_RATE
    out = np.zeros(n_samples, dtype=np.complex128)
    for k, (dibits, amp, cfo) in channels.items():
        f = (k if k < S.PFB_CHANNELS // 2 else k - S.PFB_CHANNELS) * 12500.0 + cfo
        iq = modulate_iq(dibits, fs, snr_db=None, cfo_hz=f, seed=seed + 7 * k, amplitude=amp).astype(np.complex128)
        m = min(n_samples, len(iq))
        out[:m] += iq[:m]
    if noise_db is not None:
        rng = np.random.default_rng(seed + 99)
        sigma = np.sqrt(10 ** (noise_db / 10) / 2)
        out += sigma * (rng.standard_normal(n_samples) + 1j * rng.standard_normal(n_samples))
    return out.astype(np.complex64)


def iq_to_u8(iq: np.ndarray) -> np.ndarray:
    """RTL-SDR style interleaved unsigned bytes (src/sdr.rs:25-33, src/demod.rs:72-84)."""
    v = np.empty(2 * len(iq), dtype=np.float32)
    v[0::2] = iq.real
    v[1::2] = iq.imag
    return np.clip(np.rint(v * 127.5 + 127.5), 0, 255).astype(np.uint8)


def expected_events(stream: Stream):
    """[(kind, payload)] in transmit order."""
    return [(k, p) for k, p, _ in sorted(stream.events, key=lambda e: e[2])]
