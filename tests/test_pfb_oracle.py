"""CPU checks of the channelizer restatement (oracle/pfb_oracle.py) and of the generated prototype taps."""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "spec"))

import p25_spec as S  # noqa: E402
from oracle import pfb_oracle as po  # noqa: E402
from tools import p25tx as tx  # noqa: E402


def test_polyphase_form_equals_definition():
    rng = np.random.default_rng(5)
    x = (rng.standard_normal(40000) + 1j * rng.standard_normal(40000)).astype(np.complex64)
    y = po.channelize(x, 0, 90)
    for k in (0, 1, 2, 383, 767, 768, 769, 1000, 1535):
        ms = np.array([0, 1, 14, 15, 16, 40, 89])
        d = po.channel_direct(x, k, ms)
        assert np.max(np.abs(y[ms, k] - d)) < 1e-9 * max(1.0, np.max(np.abs(d)))


def test_offset_start_matches_whole_stream():
    rng = np.random.default_rng(6)
    x = (rng.standard_normal(30000) + 1j * rng.standard_normal(30000)).astype(np.complex64)
    whole = po.channelize(x, 0, 70)
    part = po.channelize(x, 33, 20)
    assert np.allclose(whole[33:53], part, rtol=0, atol=1e-12)


def test_channel_filter_folds_into_the_prototype():
    """The identity the round-2 kernels rest on: the reference's 41-tap channel filter down every channel column equals
    ONE polyphase pass with the equivalent prototype hp (*) upsample(hc, 400) (15 taps per branch instead of 4)."""
    rng = np.random.default_rng(8)
    x = (rng.standard_normal(400 * 70) + 1j * rng.standard_normal(400 * 70)).astype(np.complex64)
    two_step = po.channel_filter(po.channelize(x))
    fused = po.channelize_fused(x)
    assert np.max(np.abs(two_step - fused)) < 1e-6 * np.max(np.abs(two_step))        # f32 rounding of the fused taps only
    heq = po.equivalent_prototype()
    assert len(heq) == 6144 + 400 * 40 and abs(heq.sum() - 1.0) < 1e-6 and -(-len(heq) // 1536) == 15


def test_prototype_filter_meets_its_spec():
    h = S.taps_pfb().astype(np.float64)
    assert len(h) == S.PFB_TAPS_LEN == 6144 and abs(h.sum() - 1.0) < 1e-6
    H = np.abs(np.fft.rfft(h, 1 << 20))
    f = np.fft.rfftfreq(1 << 20, 1.0 / S.PFB_SAMPLE_RATE)
    assert np.max(np.abs(20 * np.log10(H[f <= 6250.0]))) < 0.05            # flat over a 12.5 kHz channel
    assert np.max(20 * np.log10(H[f >= 24000.0] + 1e-30)) < -75.0          # nothing folds back at 48 kS/s


def test_generated_header_holds_these_taps():
    txt = open(os.path.join(ROOT, "p25rx_b200", "csrc", "p25_pfb_taps.h")).read()
    body = txt[txt.index("P25_TAPS_PFB_H"):]
    vals = np.array([float.fromhex(v) for v in re.findall(r"(-?0x[0-9a-fA-F.]+p[-+]?\d+)f", body)], dtype=np.float32)
    assert len(vals) == S.PFB_TAPS_LEN and (vals == S.taps_pfb()).all()
    assert f"#define P25_PFB_N {S.PFB_CHANNELS}" in txt and f"#define P25_PFB_M {S.PFB_DECIM}" in txt


def test_channel_carries_a_decodable_control_channel():
    """A transmitter 100 channels up comes out of channel 100 and decodes through the reference's 48 kHz chain."""
    from oracle import pyoracle as o
    st = tx.control_channel(7, 2)
    n = (len(st.dibits) * 10 + 200) * S.PFB_DECIM
    cap = tx.wideband_capture({100: (st.dibits, 0.05, 30.0)}, n, noise_db=-60.0, seed=1)
    y = po.channelize(cap)
    o.lib().p25o_set_always_correlate(0)
    bb = o.DemodChain(o.FMT_CF32, 2).feed(y[:, 100].astype(np.complex64))
    ev = o.MessageReceiver().feed(bb)
    tsbk = [bytes(e["payload"][:12]) for e in ev if e["kind"] == 7]
    assert len(tsbk) == 6
    for pl in tsbk:
        assert S.crc_ccitt_p25(pl[:10]) == (pl[10] << 8 | pl[11])
    quiet = o.MessageReceiver().feed(o.DemodChain(o.FMT_CF32, 2).feed(y[:, 900].astype(np.complex64)))
    assert len(quiet) == 0


def test_kernel_dataflow_equals_the_definition():
    """spec/pfb_dataflow.py restates the index arithmetic of p25_pfbc_kernel (register windows per residue class that
    advance a warp at a time, tap rows by distance to the newest sample, parity split over the two CTAs of a cluster, 8 x 8 x 12 Stockham passes, radix-2 combine, warm-up from the
    carried tail).  Three unequal chunks, odd lengths: every output time equals the fused float64 oracle."""
    import pfb_dataflow as df
    src = open(os.path.join(ROOT, "p25rx_b200", "csrc", "pfb.cu")).read()
    for name in ("HTX", "WARM", "TB", "NTH", "WS", "EARLY"):       # the model and the kernel agree on their constants
        assert int(re.search(rf"constexpr int {name} = (\d+);", src).group(1)) == getattr(df, name), name
    rng = np.random.default_rng(0)
    v = rng.standard_normal(768) + 1j * rng.standard_normal(768)
    assert np.max(np.abs(df.idft768(v) - np.fft.ifft(v) * 768)) < 1e-10
    # the padded buffer layouts: injective, inside the buffer, and free of bank conflicts on both sides of each pass
    # (a half-warp's sixteen 8-byte accesses must fall into sixteen different bank pairs)
    idx = np.arange(768)
    for f in (df.skew_a, df.skew_b):
        assert len(set(f(idx))) == 768 and f(idx).max() < 768 + 768 // 8
    banks = lambda pos: len(set(int(x) % 16 for x in pos))
    for j0 in range(0, 96, 16):
        j = np.arange(j0, j0 + 16)
        assert all(banks(df.skew_a(8 * j + q)) == 16 for q in range(8))                       # pass A writes
        assert all(banks(df.skew_a(j + 96 * r)) == 16 for r in range(8))                      # pass B reads
        assert all(banks(df.skew_b((j >> 3) * 64 + (j & 7) + 8 * q)) == 16 for q in range(8))  # pass B writes
    for j0 in range(0, 64, 16):
        j = np.arange(j0, j0 + 16)
        assert all(banks(df.skew_b(j + 64 * r)) == 16 for r in range(12))                     # pass C reads
    heq = po.equivalent_prototype()
    x = rng.standard_normal(36 * 400 + 123) + 1j * rng.standard_normal(36 * 400 + 123)
    ref = po.channelize_fused(x)
    tail, pos = np.zeros(df.HTX, dtype=np.complex128), 0
    for n in (5001, 6777, len(x) - 5001 - 6777):
        m0 = max(0, -(-(pos - 399) // 400))                # first output whose newest input lies in this chunk
        n_out = (pos + n - 400) // 400 - m0 + 1
        y = df.Run(heq, tail, x[pos:pos + n], pos, m0).spectra(0, n_out)[df.WARM:]
        assert np.max(np.abs(y - ref[m0:m0 + n_out])) < 1e-12 * np.max(np.abs(ref))
        tail = np.concatenate([tail, x[pos:pos + n]])[-df.HTX:]
        pos += n
