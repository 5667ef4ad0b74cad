"""Normative constants and code definitions for the P25 Phase 1 baseband hot path.

This module is the single source of truth for every table used by
  * the CPU oracle            (oracle/p25_tables.h, generated)
  * the CUDA product library  (p25rx_b200/csrc/p25_tables.cuh, generated)
  * the synthetic transmitter (tools/p25tx.py, imports this module directly)

Provenance tags (same legend as SURVEY.md):
  [REF]    visible in /root/reference (file:line cited)
  [STD]    TIA-102.BAAA-A Common Air Interface, recalled from memory; every such
           constant is cross-checked by an algebraic self-test in
           tests/test_spec_selfcheck.py (minimum distance, divisibility, ...)
  [RECALL] recollection of an un-vendored crate (kchmck/p25.rs @a96c564, ...)
  [BUILD]  chosen by this build because the reference delegates it to a crate
           whose source is not available here (p25_filts taps, sync rule, ...)

Nothing in here is copied from the reference tree.
"""
from __future__ import annotations

import numpy as np

# ----------------------------------------------------------------------------
# Rates and chunk sizes  [REF src/consts.rs:6-13, src/demod.rs:50-54]
# ----------------------------------------------------------------------------
SDR_SAMPLE_RATE = 240_000        # [REF src/consts.rs:11]
BASEBAND_SAMPLE_RATE = 48_000    # [REF src/consts.rs:13]
BUF_BYTES = 32_768               # [REF src/consts.rs:6]
BUF_SAMPLES = BUF_BYTES // 2     # [REF src/consts.rs:8]
DECIM_NATIVE = 5                 # [REF src/demod.rs:50]
BOXCAR_LEN = 10                  # [REF src/demod.rs:52]
FM_DEVIATION_HZ = 5000           # [REF src/demod.rs:54]
FM_GAIN = np.float32(BASEBAND_SAMPLE_RATE / (2.0 * np.pi * FM_DEVIATION_HZ))  # [RECALL demod_fm]
WIDE_SAMPLE_RATE = 2_400_000     # BASELINE.json configs[0]; declared extension
DECIM_FRONT = 10                 # 2.4 MS/s -> 240 kS/s front stage [BUILD]

SYMBOL_RATE = 4800               # [STD]
SPS = BASEBAND_SAMPLE_RATE // SYMBOL_RATE   # 10 samples / symbol

# ----------------------------------------------------------------------------
# FIR taps [BUILD]: p25_filts' DecimFir / BandpassFir values are not available.
# Designed with scipy (kaiser window), rounded to f32, DC gain 1.
# ----------------------------------------------------------------------------
TAPS_FRONT_LEN = 50   # /10 stage at 2.4 MS/s
TAPS_DECIM_LEN = 25   # /5 stage at 240 kS/s   (role of p25_filts::DecimFir)
TAPS_CHAN_LEN = 41    # channel select at 48 kS/s (role of p25_filts::BandpassFir)


def _kaiser_lowpass(ntaps: int, cutoff_hz: float, fs: float, atten_db: float) -> np.ndarray:
    from scipy import signal
    beta = signal.kaiser_beta(atten_db)
    h = signal.firwin(ntaps, cutoff_hz, window=("kaiser", beta), fs=fs)
    h = h / h.sum()
    return h.astype(np.float32)


def taps_front() -> np.ndarray:
    return _kaiser_lowpass(TAPS_FRONT_LEN, 104_000.0, WIDE_SAMPLE_RATE, 60.0)


def taps_decim() -> np.ndarray:
    return _kaiser_lowpass(TAPS_DECIM_LEN, 24_000.0, SDR_SAMPLE_RATE, 60.0)


def taps_chan() -> np.ndarray:
    return _kaiser_lowpass(TAPS_CHAN_LEN, 6_250.0, BASEBAND_SAMPLE_RATE, 50.0)


# ----------------------------------------------------------------------------
# Wideband polyphase channelizer [BUILD] (BASELINE.json configs[2]; no reference stage):
# a 19.2 MS/s cf32 capture is split into 1,536 channels 12.5 kHz apart, each delivered at 48 kS/s
# (decimation 400, oversampled 3.84x).  Channel k is DEFINED as mix-down by k * 12.5 kHz, the prototype
# low-pass below, and decimation by 400 with the newest input of output m at index 400 m + 399:
#     y_k[m] = sum_i h[i] * x[400 m + 399 - i] * exp(-2j pi k (400 m + 399 - i) / 1536)
# followed by the reference's own 48 kHz stages (channel-select FIR, discriminator, boxcar).
# ----------------------------------------------------------------------------
PFB_SAMPLE_RATE = 19_200_000
PFB_CHANNELS = 1536
PFB_DECIM = 400
PFB_TAPS_PER_BRANCH = 4
PFB_TAPS_LEN = PFB_CHANNELS * PFB_TAPS_PER_BRANCH


def taps_pfb() -> np.ndarray:
    """Prototype low-pass: 16 kHz cutoff, 80 dB Kaiser; pass band flat to +-8 kHz, stop band from 24 kHz."""
    return _kaiser_lowpass(PFB_TAPS_LEN, 16_000.0, PFB_SAMPLE_RATE, 80.0)


def iq_lut() -> np.ndarray:
    """u8 -> f32 mapping of rtlsdr_iq [RECALL, scale unpinned]: (b - 127.5) / 127.5."""
    b = np.arange(256, dtype=np.float32)
    return ((b - np.float32(127.5)) / np.float32(127.5)).astype(np.float32)


# ----------------------------------------------------------------------------
# Symbols, frame sync, status symbols  [STD]
# ----------------------------------------------------------------------------
FRAME_SYNC = 0x5575F5FF77FF            # 48 bits, 24 dibits
FS_DIBITS = 24
NID_DIBITS = 32
STATUS_PERIOD = 36                     # every 36th dibit of a data unit is a status symbol
DIBIT_TO_SYMBOL = {0b01: +3, 0b00: +1, 0b10: -1, 0b11: -3}
SYMBOL_DEVIATION_HZ = 600.0            # per unit symbol: +-1800 / +-600 Hz


def frame_sync_dibits() -> np.ndarray:
    return np.array([(FRAME_SYNC >> (46 - 2 * i)) & 3 for i in range(24)], dtype=np.uint8)


def frame_sync_symbols() -> np.ndarray:
    return np.array([DIBIT_TO_SYMBOL[int(d)] for d in frame_sync_dibits()], dtype=np.int32)


# Data unit IDs [STD]; names follow src/policy.rs:99-111
DUID_HDU, DUID_TDU, DUID_LDU1, DUID_TSDU, DUID_LDU2, DUID_PDU, DUID_TDULC = 0x0, 0x3, 0x5, 0x7, 0xA, 0xC, 0xF

# ----------------------------------------------------------------------------
# Sync correlator [BUILD, shaped after RECALL of p25.rs baseband/sync.rs]
# fingerprint = ideal post-integrator waveform of the 24 sync symbols, sampled at
# 10 samples/symbol from the centre of symbol 0 to the centre of symbol 23.
# ----------------------------------------------------------------------------
FP_LEN = (FS_DIBITS - 1) * SPS + 1     # 231
RC_ALPHA = 0.2                         # [STD] C4FM Nyquist raised cosine roll-off
SYNC_RHO = 0.65                        # normalised correlation needed to call a peak


def raised_cosine(t: np.ndarray, alpha: float = RC_ALPHA) -> np.ndarray:
    """Raised-cosine impulse response, t in symbol periods."""
    t = np.asarray(t, dtype=np.float64)
    out = np.sinc(t)
    den = 1.0 - (2.0 * alpha * t) ** 2
    sing = np.abs(den) < 1e-9
    safe = np.where(sing, 1.0, den)
    out = out * np.cos(np.pi * alpha * t) / safe
    out = np.where(sing, (np.pi / 4.0) * np.sinc(1.0 / (2.0 * alpha)), out)
    return out


def sync_fingerprint() -> np.ndarray:
    syms = frame_sync_symbols().astype(np.float64)
    k = np.arange(FP_LEN, dtype=np.float64)
    fp = np.zeros(FP_LEN)
    for i, s in enumerate(syms):
        fp += s * raised_cosine((k - SPS * i) / SPS)
    fp /= 3.0                                     # outer symbols -> +-1
    return fp.astype(np.float32)


def sync_rho2_efp() -> np.float32:
    fp = sync_fingerprint().astype(np.float64)
    return np.float32(SYNC_RHO * SYNC_RHO * float(np.sum(fp * fp)))


# ----------------------------------------------------------------------------
# GF(2^6), primitive polynomial x^6 + x + 1  [STD]
# ----------------------------------------------------------------------------
GF_POLY = 0x43


def gf_tables():
    exp = np.zeros(128, dtype=np.uint8)
    log = np.zeros(64, dtype=np.uint8)
    x = 1
    for i in range(63):
        exp[i] = x
        log[x] = i
        x <<= 1
        if x & 0x40:
            x ^= GF_POLY
    for i in range(63, 128):
        exp[i] = exp[i - 63]
    return exp, log


GF_EXP, GF_LOG = gf_tables()


def gf_mul(a: int, b: int) -> int:
    if a == 0 or b == 0:
        return 0
    return int(GF_EXP[int(GF_LOG[a]) + int(GF_LOG[b])])


def gf_inv(a: int) -> int:
    return int(GF_EXP[(63 - int(GF_LOG[a])) % 63])


# ----------------------------------------------------------------------------
# BCH(63,16,23) for the NID  [STD generator, octal 6331141367235453]
# ----------------------------------------------------------------------------
BCH_GEN = 0o6331141367235453           # degree 47
BCH_T = 11


def _polymod2(a: int, g: int) -> int:
    dg = g.bit_length() - 1
    while a.bit_length() - 1 >= dg and a:
        a ^= g << (a.bit_length() - 1 - dg)
    return a


def bch_encode(data16: int) -> int:
    """16-bit (NAC<<4 | DUID) -> 63-bit systematic codeword."""
    m = (data16 & 0xFFFF) << 47
    return m | _polymod2(m, BCH_GEN)


def nid_encode(nac: int, duid: int) -> int:
    """64-bit NID: 63-bit BCH codeword followed by one overall parity bit [STD]."""
    cw = bch_encode(((nac & 0xFFF) << 4) | (duid & 0xF))
    return (cw << 1) | (bin(cw).count("1") & 1)


def bch_derived_generator() -> int:
    """lcm of the minimal polynomials of alpha^1..alpha^22 - used by the self-test."""
    seen = set()
    g = 1
    for i in range(1, 2 * BCH_T + 1):
        # conjugacy class of alpha^i
        cls = []
        e = i % 63
        while e not in cls:
            cls.append(e)
            e = (e * 2) % 63
        key = min(cls)
        if key in seen:
            continue
        seen.add(key)
        # minimal polynomial prod (x + alpha^e), coefficients in GF(64) -> ends in GF(2)
        poly = [1]
        for e in cls:
            r = int(GF_EXP[e])
            new = [0] * (len(poly) + 1)
            for d, c in enumerate(poly):
                new[d + 1] ^= c
                new[d] ^= gf_mul(c, r)
            poly = new
        assert all(c in (0, 1) for c in poly)
        mp = sum(c << d for d, c in enumerate(poly))
        # GF(2) polynomial product
        prod = 0
        for d in range(mp.bit_length()):
            if (mp >> d) & 1:
                prod ^= g << d
        g = prod
    return g


# ----------------------------------------------------------------------------
# Golay codes  [STD g(x) = 0xC75]
# ----------------------------------------------------------------------------
GOLAY_GEN = 0xC75


def golay23_encode(data12: int) -> int:
    m = (data12 & 0xFFF) << 11
    return m | _polymod2(m, GOLAY_GEN)


def golay24_encode(data12: int) -> int:
    cw = golay23_encode(data12)
    return (cw << 1) | (bin(cw).count("1") & 1)


def golay18_encode(data6: int) -> int:
    return golay24_encode(data6 & 0x3F) & 0x3FFFF


def golay23_syndrome_table() -> np.ndarray:
    """syndrome (11 bit) -> 23-bit error pattern of weight <= 3 (perfect code)."""
    import itertools
    tab = np.zeros(2048, dtype=np.uint32)
    filled = np.zeros(2048, dtype=bool)
    filled[0] = True
    for w in (1, 2, 3):
        for pos in itertools.combinations(range(23), w):
            e = 0
            for p in pos:
                e |= 1 << p
            s = _polymod2(e, GOLAY_GEN)
            assert not filled[s]
            tab[s] = e
            filled[s] = True
    assert filled.all()
    return tab


# ----------------------------------------------------------------------------
# Hamming codes  [STD generator matrices; data bits then parity bits]
# Parity columns (one 4-bit value per data bit, MSB data bit first).
# ----------------------------------------------------------------------------
HAMMING15_COLS = [15, 14, 13, 12, 11, 10, 9, 7, 6, 5, 3]   # (15,11,3)
HAMMING10_COLS = [14, 13, 11, 7, 3, 12]                    # (10,6,3)


def hamming15_encode(data11: int) -> int:
    p = 0
    for i, c in enumerate(HAMMING15_COLS):
        if (data11 >> (10 - i)) & 1:
            p ^= c
    return ((data11 & 0x7FF) << 4) | p


def hamming10_encode(data6: int) -> int:
    p = 0
    for i, c in enumerate(HAMMING10_COLS):
        if (data6 >> (5 - i)) & 1:
            p ^= c
    return ((data6 & 0x3F) << 4) | p


def hamming15_syndrome_table() -> np.ndarray:
    """syndrome (4 bit) -> bit mask to flip in the 15-bit word (0 for syndrome 0)."""
    tab = np.zeros(16, dtype=np.uint16)
    for i, c in enumerate(HAMMING15_COLS):
        tab[c] = 1 << (14 - i)
    for j in range(4):
        tab[1 << j] = 1 << j
    return tab


def hamming10_syndrome_table() -> np.ndarray:
    """syndrome -> flip mask in the 10-bit word; 0xFFFF marks an unrecoverable syndrome."""
    tab = np.full(16, 0xFFFF, dtype=np.uint16)
    tab[0] = 0
    for i, c in enumerate(HAMMING10_COLS):
        tab[c] = 1 << (9 - i)
    for j in range(4):
        tab[1 << j] = 1 << j
    return tab


# ----------------------------------------------------------------------------
# Cyclic (16,8,5) for low speed data  [STD: shortened (17,9) code; g(x) recalled]
# ----------------------------------------------------------------------------
CYCLIC_GEN = 0x139      # x^8 + x^5 + x^4 + x^3 + 1, a factor of x^17 + 1


def cyclic16_encode(data8: int) -> int:
    m = (data8 & 0xFF) << 8
    return m | _polymod2(m, CYCLIC_GEN)


def cyclic16_syndrome_table() -> np.ndarray:
    """syndrome (8 bit) -> 16-bit error pattern of weight <= 2, 0xFFFF = unrecoverable."""
    import itertools
    tab = np.full(256, 0xFFFF, dtype=np.uint16)
    tab[0] = 0
    for w in (1, 2):
        for pos in itertools.combinations(range(16), w):
            e = 0
            for p in pos:
                e |= 1 << p
            s = _polymod2(e, CYCLIC_GEN)
            assert tab[s] == 0xFFFF, "cyclic code distance < 5"
            tab[s] = e
    return tab


# ----------------------------------------------------------------------------
# Reed-Solomon over GF(2^6)  [STD: RS(24,12,13), RS(24,16,9), RS(36,20,17)]
# systematic, generator roots alpha^1 .. alpha^(n-k); symbol 0 = highest degree
# ----------------------------------------------------------------------------
RS_SHORT = (24, 12)
RS_MED = (24, 16)
RS_LONG = (36, 20)


def rs_generator(nroots: int) -> list[int]:
    """coefficients, index = degree."""
    g = [1]
    for i in range(1, nroots + 1):
        r = int(GF_EXP[i])
        new = [0] * (len(g) + 1)
        for d, c in enumerate(g):
            new[d + 1] ^= c
            new[d] ^= gf_mul(c, r)
        g = new
    return g


def rs_encode(data: list[int], n: int, k: int) -> list[int]:
    assert len(data) == k
    nroots = n - k
    g = rs_generator(nroots)
    rem = [0] * nroots          # rem[0] = highest degree of the remainder
    for d in data:
        fb = d ^ rem[0]
        rem = rem[1:] + [0]
        if fb:
            for j in range(nroots):
                rem[j] ^= gf_mul(fb, g[nroots - 1 - j])
    return list(data) + rem


# ----------------------------------------------------------------------------
# Trellis codes and the data interleaver  [STD]
# ----------------------------------------------------------------------------
TRELLIS_HALF = [            # [state][input dibit] -> constellation point
    [0, 15, 12, 3],
    [4, 11, 8, 7],
    [13, 2, 1, 14],
    [9, 6, 5, 10],
]
TRELLIS_3_4 = [             # [state][input tribit] -> constellation point
    [0, 8, 4, 12, 2, 10, 6, 14],
    [4, 12, 2, 10, 6, 14, 0, 8],
    [1, 9, 5, 13, 3, 11, 7, 15],
    [5, 13, 3, 11, 7, 15, 1, 9],
    [3, 11, 7, 15, 1, 9, 5, 13],
    [7, 15, 1, 9, 5, 13, 3, 11],
    [2, 10, 6, 14, 0, 8, 4, 12],
    [6, 14, 0, 8, 4, 12, 2, 10],
]
# constellation point -> (first dibit << 2 | second dibit)
CONSTELLATION = [0x2, 0xA, 0x7, 0xF, 0xE, 0x6, 0xB, 0x3, 0xD, 0x5, 0x8, 0x0, 0x1, 0x9, 0x4, 0xC]

TSBK_DIBITS = 98
TSBK_BYTES = 12


def interleave_perm() -> list[int]:
    """perm[i] = transmitted 4-bit-symbol slot that carries trellis symbol i.

    [STD] interleave table: trellis symbols 0,1,2,3,4,... go to slots 0,13,25,37,1,...
    """
    order = []
    for r in range(12):
        order += [r, 13 + r, 25 + r, 37 + r]
    order.append(12)
    assert sorted(order) == list(range(49))
    return order


def trellis_half_encode(dibits48: list[int]) -> list[int]:
    """48 dibits + flush -> 49 constellation points."""
    state = 0
    out = []
    for d in list(dibits48) + [0]:
        out.append(TRELLIS_HALF[state][d])
        state = d
    return out


def tsbk_block_dibits(payload12: bytes) -> np.ndarray:
    """12-byte TSBK -> 98 transmitted dibits (trellis + interleave)."""
    assert len(payload12) == 12
    bits = np.unpackbits(np.frombuffer(bytes(payload12), dtype=np.uint8))
    dibits = [(int(bits[2 * i]) << 1) | int(bits[2 * i + 1]) for i in range(48)]
    return _points_to_dibits(trellis_half_encode(dibits))


def trellis_34_encode(tribits48: list[int]) -> list[int]:
    """48 tribits + flush -> 49 constellation points (3/4-rate, 8 states; next state = input tribit) [STD]."""
    state = 0
    out = []
    for t in list(tribits48) + [0]:
        out.append(TRELLIS_3_4[state][t])
        state = t
    return out


def _points_to_dibits(pts: list[int]) -> np.ndarray:
    perm = interleave_perm()
    slots = [0] * 49
    for i, p in enumerate(pts):
        slots[perm[i]] = CONSTELLATION[p]
    out = np.zeros(98, dtype=np.uint8)
    for s, v in enumerate(slots):
        out[2 * s] = (v >> 2) & 3
        out[2 * s + 1] = v & 3
    return out


PDU_BLOCK34_BYTES = 18


def pdu_block34_dibits(payload18: bytes) -> np.ndarray:
    """18-byte confirmed-data block -> 98 transmitted dibits (3/4-rate trellis + the data interleaver) [STD]."""
    assert len(payload18) == PDU_BLOCK34_BYTES
    bits = np.unpackbits(np.frombuffer(bytes(payload18), dtype=np.uint8))
    tribits = [(int(bits[3 * i]) << 2) | (int(bits[3 * i + 1]) << 1) | int(bits[3 * i + 2]) for i in range(48)]
    return _points_to_dibits(trellis_34_encode(tribits))


# Packet data unit header [STD layout, RECALL]: octet 0 = 0 | A/N | I/O | format(5); octet 1 = 1 1 | SAP(6); octet 2 MFID;
# octets 3-5 logical link id; octet 6 = FMF | blocks to follow (7); octet 7 pad count; octet 8 Syn | N(S) | FSNF;
# octet 9 data header offset; octets 10-11 CRC-CCITT of octets 0-9.  Format 0x16 = confirmed data (3/4-rate blocks of
# 18 octets), 0x15 = unconfirmed data (1/2-rate blocks of 12 octets).
PDU_FORMAT_CONFIRMED = 0x16
PDU_FORMAT_UNCONFIRMED = 0x15
PDU_MAX_BLOCKS = 127


def pdu_header(fmt: int, blocks_to_follow: int, sap: int = 0x04, mfid: int = 0, llid: int = 0x123456, pad: int = 0,
               bad_crc: bool = False) -> bytes:
    head = bytes([0x40 | (fmt & 0x1F), 0xC0 | (sap & 0x3F), mfid & 0xFF, (llid >> 16) & 0xFF, (llid >> 8) & 0xFF, llid & 0xFF,
                  0x80 | (blocks_to_follow & 0x7F), pad & 0x1F, 0x00, 0x00])
    crc = crc_ccitt_p25(head)
    if bad_crc:
        crc ^= 0x0440
    return head + bytes([crc >> 8, crc & 0xFF])


def crc9_p25(bits) -> int:
    """CRC-9 of a confirmed data block [RECALL: g = x^9+x^6+x^4+x^3+1, inverted]; carried, never checked by the
    MessageReceiver surface (no event holds packet data)."""
    crc = 0
    for b in bits:
        crc = ((crc << 1) | int(b)) & 0x3FF
        if crc & 0x200:
            crc ^= 0x259
    for _ in range(9):
        crc = (crc << 1) & 0x3FF
        if crc & 0x200:
            crc ^= 0x259
    return (crc ^ 0x1FF) & 0x1FF


# Path-metric bounds above which a trellis block is reported undecodable [BUILD]: random 98-dibit blocks decode to
# metric >= 22 (1/2 rate) and >= 8 (3/4 rate) (tests/test_oracle_fec.py).
VITERBI_MAX_FIX = 18
VITERBI34_MAX_FIX = 6


def crc_ccitt_p25(data: bytes) -> int:
    """CRC-CCITT of TSBKs [STD]: g = x^16+x^12+x^5+1, zero init, result inverted."""
    crc = 0
    for byte in data:
        crc ^= byte << 8
        for _ in range(8):
            crc = ((crc << 1) ^ 0x1021) & 0xFFFF if crc & 0x8000 else (crc << 1) & 0xFFFF
    return crc ^ 0xFFFF


# ----------------------------------------------------------------------------
# IMBE voice frame layout  [STD for code sizes, PN generator; interleave schedule:
# first 28 symbols as recalled from the vocoder annex, remainder BUILD-defined by
# the same round-robin rule]
# ----------------------------------------------------------------------------
IMBE_CW_BITS = [23, 23, 23, 23, 15, 15, 15, 7]
IMBE_U_BITS = [12, 12, 12, 12, 11, 11, 11, 7]
IMBE_DIBITS = 72


def imbe_schedule() -> np.ndarray:
    """[144][2] = (codeword index, bit index (0 = LSB)) for each transmitted bit."""
    left = list(IMBE_CW_BITS)
    order = []
    while sum(left):
        for c in range(8):
            if left[c]:
                left[c] -= 1
                order.append((c, left[c]))
    assert len(order) == 144
    return np.array(order, dtype=np.uint8)


def imbe_pn_masks(u0: int) -> list[int]:
    """scrambling masks m0..m7 for codewords c0..c7 (m0 = m7 = 0) [STD]."""
    p = (16 * (u0 & 0xFFF)) & 0xFFFF
    masks = [0] * 8
    for c in range(1, 7):
        m = 0
        for _ in range(IMBE_CW_BITS[c]):
            p = (173 * p + 13849) & 0xFFFF
            m = (m << 1) | (p >> 15)
        masks[c] = m
    return masks


def imbe_encode(u: list[int]) -> np.ndarray:
    """u0..u7 -> 72 transmitted dibits."""
    masks = imbe_pn_masks(u[0])
    cw = []
    for c in range(8):
        if c < 4:
            w = golay23_encode(u[c])
        elif c < 7:
            w = hamming15_encode(u[c])
        else:
            w = u[c] & 0x7F
        cw.append(w ^ masks[c])
    sched = imbe_schedule()
    bits = [(int(cw[int(c)]) >> int(b)) & 1 for c, b in sched]
    return np.array([(bits[2 * i] << 1) | bits[2 * i + 1] for i in range(72)], dtype=np.uint8)


# LDU payload layout in data dibits after the NID [STD]
LDU_LAYOUT = [("vf", 72), ("vf", 72), ("lc", 20), ("vf", 72), ("lc", 20), ("vf", 72), ("lc", 20),
              ("vf", 72), ("lc", 20), ("vf", 72), ("lc", 20), ("vf", 72), ("lc", 20), ("vf", 72),
              ("lsd", 16), ("vf", 72)]
LDU_DIBITS = sum(n for _, n in LDU_LAYOUT)      # 784
HDU_DIBITS = 324
TDULC_DIBITS = 144

# Error codes carried by MessageEvent::Error [BUILD numbering; families after src/hub.rs:557-572]
ERR_BCH, ERR_RS, ERR_VITERBI_DIBIT, ERR_VITERBI_TRIBIT, ERR_UNKNOWN_NID = 1, 2, 3, 4, 5

# Stats families in serialisation order [REF src/hub.rs:557-572]
STATS_FAMILIES = ["bch", "cyclic", "golayStd", "golayExt", "golayShort", "hammingStd",
                  "hammingShort", "rsShort", "rsMed", "rsLong", "viterbiDibit", "viterbiTribit"]
# symbols per word per family (CodeStats.size) [RECALL]
STATS_SIZE = [63, 16, 23, 24, 18, 15, 10, 24, 24, 36, 98, 98]
