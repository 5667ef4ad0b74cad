"""CPU check of the integer-tensor-pipe decimator's decomposition (spec/w5i_dataflow.py) against the float64 FIR:
fragment layouts, banded tap matrix, row permutation, limbs, accumulator start values, magic-number conversion, and the
shared-memory bank mapping of its loads and stores."""
import numpy as np
import pytest

import p25_spec as spec
import w5i_dataflow as df


def _taps():
    return spec.taps_decim().astype(np.float32)


def test_limbs_reconstruct_the_taps():
    taps = _taps()
    lb, tsum = df.limbs(taps)
    t = lb[:, 0] + 256 * lb[:, 1] + 65536 * lb[:, 2]
    assert np.all(np.abs(lb) <= 128) and np.all(lb >= -128) and np.all(lb <= 127)
    assert np.max(np.abs(t / 2.0 ** df.SCALE_LOG2 - taps.astype(np.float64))) <= 2.0 ** -(df.SCALE_LOG2 + 1)
    assert tsum == int(t.sum())


@pytest.mark.parametrize("skew", [0, 4, 5, 6, 7])     # the values l_base & 7 takes (ht - 20 + 0..4)
def test_iteration_equals_the_fir(skew):
    rng = np.random.default_rng(skew)
    taps = _taps()
    xs = rng.integers(0, 256, size=2688, dtype=np.int64)
    yf, load_banks, store_words = df.decimate_iteration(xs, skew, taps)
    lb, _ = df.limbs(taps)
    tq = (lb[:, 0] + 256 * lb[:, 1] + 65536 * lb[:, 2]).astype(np.int64)
    xi = xs[0::2] - 128
    xq = xs[1::2] - 128
    for o in range(df.NOUT):
        w = slice(skew + 5 * o, skew + 5 * o + df.NTAPS)
        # y[o] = sum_k h[k] x[5 o + 24 - k]
        ei = int(np.dot(tq[::-1], xi[w]))
        eq = int(np.dot(tq[::-1], xq[w]))
        assert abs(float(yf[o, 0]) - ei) <= abs(ei) * 2.0 ** -23 + 1 and abs(float(yf[o, 1]) - eq) <= abs(eq) * 2.0 ** -23 + 1
    # float taps, float64 arithmetic: the integer path is the same filter to 2^-26 per tap
    o = 77
    ref = np.dot(taps[::-1].astype(np.float64), xi[skew + 5 * o: skew + 5 * o + df.NTAPS]) * 2.0 ** df.SCALE_LOG2
    assert abs(float(yf[o, 0]) - ref) <= 128 * df.NTAPS * 0.5 + abs(ref) * 2.0 ** -23
    for banks in load_banks:                       # 4-byte loads: 32 lanes, 32 different banks
        assert len(set(banks.tolist())) == 32
    for words in store_words:                      # 8-byte stores: each half-warp covers 32 different banks
        for half in (words[:16], words[16:]):
            banks = np.concatenate([half % 32, (half + 1) % 32])
            assert len(set(banks.tolist())) == 32


@pytest.mark.parametrize("skew", [0, 3, 6])
def test_combined_50_iteration_equals_the_two_stage_chain(skew):
    """w50i: the 290-tap integer FIR on the bytes equals front (/10, 50 taps) then decimator (/5, 25 taps) in float64."""
    rng = np.random.default_rng(10 + skew)
    hf, hd = spec.taps_front().astype(np.float32), spec.taps_decim().astype(np.float32)
    g = df.combined_taps(hf, hd)
    assert len(g) == df.G50 and abs(g.sum() - hf.astype(np.float64).sum() * hd.astype(np.float64).sum()) < 1e-12
    xs = rng.integers(0, 256, size=13312, dtype=np.int64)
    yf, load_banks, stores = df.decimate50_iteration(xs, skew, g)
    x = (xs[0::2] - 128) + 1j * (xs[1::2] - 128)
    # two-stage definition on the same samples: f[n] = sum_i hf[i] X[10 n + 9 - i], yd[t] = sum_k hd[k] f[5 t + 4 - k],
    # X relative to the block base = window start + 240
    base = skew + 240
    for o in (0, 1, 63, 64, 100, 127):
        acc = 0.0
        for k in range(25):
            n = 5 * o + 4 - k
            idx = base + 10 * n + 9 - np.arange(50)
            acc += float(hd[k]) * np.dot(hf.astype(np.float64), x[idx])
        ref = acc * 2.0 ** df.SCALE50
        got = float(yf[o, 0]) + 1j * float(yf[o, 1])
        assert abs(got - ref) <= 128 * df.G50 * 0.5 * 1.5 + abs(ref) * 2.0 ** -22, (o, got, ref)
    for banks in load_banks:
        assert len(set(banks.tolist())) == 32
    for words in stores:
        for half in (words[:16], words[16:]):
            banks = np.concatenate([half % 32, (half + 1) % 32])
            assert len(set(banks.tolist())) == 32
