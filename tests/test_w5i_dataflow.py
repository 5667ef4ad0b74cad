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


# ------------------------------------------------------------------ the library's own tables (p25_imma_tables.h) vs the model
@pytest.fixture(scope="module")
def imma_hostcheck():
    """tests/hostcheck: the header ddc_fm.cu builds its table uploads from, compiled for the host -- test harness only."""
    import ctypes as C
    import os
    import subprocess
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostcheck")
    so = os.path.join(d, "libimma_hostcheck.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-o", so, os.path.join(d, "imma_hostcheck.cpp")])
    return C.CDLL(so)


def _pack(tab):
    """model fragments [..., lane, reg, byte] (signed) -> words [..., lane, reg] like the library's Frag entries"""
    b = (tab & 255).astype(np.uint64)
    return (b[..., 0] | (b[..., 1] << 8) | (b[..., 2] << 16) | (b[..., 3] << 24)).astype(np.uint32)


def test_library_tables_equal_the_model(imma_hostcheck):
    import ctypes as C
    taps = _taps()
    lb, tsum = df.limbs(taps)
    out = np.zeros((4, 18, 32, 2), dtype=np.uint32)
    init = np.zeros(3, dtype=np.int32)
    gain = C.c_double()
    assert imma_hostcheck.hc_imma5(out.ctypes.data_as(C.c_void_p), init.ctypes.data_as(C.c_void_p), C.byref(gain)) == 1
    for sk in range(4):
        model = _pack(df.b_table(lb, sk)).reshape(18, 32, 2)       # [nt][jj][limb] -> (nt * 3 + jj) * 3 + limb
        assert np.array_equal(out[sk], model), sk
    assert init.tolist() == [df.MAGIC - 128 * int(lb[:, l].sum()) for l in range(3)]
    assert gain.value == tsum / 2.0 ** df.SCALE_LOG2

    g = df.combined_taps(spec.taps_front().astype(np.float32), taps)
    lb50 = df.limbs50(g)
    out50 = np.zeros((2, 84, 32, 2), dtype=np.uint32)
    assert imma_hostcheck.hc_imma50(out50.ctypes.data_as(C.c_void_p), init.ctypes.data_as(C.c_void_p), C.byref(gain)) == 1
    for sk in range(2):
        model = _pack(df.b_table50(lb50, sk)).reshape(84, 32, 2)   # [ks][limb] -> ks * 3 + limb
        assert np.array_equal(out50[sk], model), sk
    assert init.tolist() == [df.MAGIC - 128 * int(lb50[:, l].sum()) for l in range(3)]
    t50 = lb50[:, 0] + 256 * lb50[:, 1] + 65536 * lb50[:, 2]
    assert gain.value == int(t50.sum()) / 2.0 ** df.SCALE50


# ------------------------------------------------------------------ slice arithmetic: every kept output sees only copied bytes
@pytest.mark.parametrize("name,D,T,NOUT,XLEN,ht,clamp", [
    ("w5i", 5, 25, 256, 1312, 1312, False),        # ddc_fm.cu: w5i::issue_slice (no clamp: a slice never starts in front of the tail)
    ("w50i", 50, 290, 128, 6648, 3264, True),      # w50i::issue_slice (the part in front of the tail is skipped)
])
def test_slices_cover_every_kept_output(name, D, T, NOUT, XLEN, ht, clamp):
    """The fast kernels copy [la, lb) of the logical input (tail ++ chunk) per warp iteration.  Restated here: for any chunk
    position and length the windows of all stored outputs, and of the last 50 outputs of a warm-up iteration (40 channel
    filter + 10 boxcar histories), lie inside the copied range -- so the bytes that are not copied (in front of the tail, past
    the chunk's end) only ever reach outputs nobody keeps."""
    rng = np.random.default_rng(5)
    cases = [(ht, ht), (ht, ht + 8), (ht + 3, 16384), (12345 * 8 + 1, 360000)]
    cases += [(int(ht + rng.integers(0, 100000)), int(8 * rng.integers(ht // 8, 6000))) for _ in range(60)]
    for a0, n in cases:
        newest0 = D - 1                                    # output m's newest input is D m + D - 1
        m0 = -((newest0 - a0) // D)                        # first m with D m + D - 1 >= a0
        m_end = -((newest0 - (a0 + n)) // D)               # first m beyond the chunk
        n_out = m_end - m0
        assert n_out > 0
        l_base = D * m0 - a0 - (T - D) + ht                # logical index of output m0's oldest input
        assert 0 <= l_base and l_base + T - 1 < ht + n and (D * m0 + D - 1 - a0) in range(D)
        lend = ht + n
        ips = -(-n_out // NOUT)
        XNEW = D * NOUT
        for it in range(-1, ips):
            l0 = l_base + XNEW * it
            la0 = l0 & ~7
            la = max(la0, 0) if clamp else la0
            lb = min(la0 + XLEN, lend)
            assert la >= 0 and lb > la and (lb - la) % 8 == 0, (name, a0, n, it)
            if it >= 0:                                    # a stored iteration: outputs o < nv
                nv = min(n_out - NOUT * it, NOUT)
                assert l0 >= la and l0 + D * (nv - 1) + T - 1 < lb, (name, a0, n, it)
            if it + 1 < ips:                               # this iteration may be the warm-up of a piece starting at it + 1
                first = l0 + D * (NOUT - 50)
                assert first >= la and l0 + D * (NOUT - 1) + T - 1 < lb, (name, a0, n, it)
