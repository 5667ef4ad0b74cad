"""Non-twin cross-check of the FEC decoders: the oracle's algebraic C++ decoders -- and the
host build of the device decoder source -- against spec/p25_refdec.py, which decodes by exhaustive nearest-code-word
search over the code books the ENCODERS generate (binary codes), by Euclid + a Vandermonde solve (Reed-Solomon) and by
a batched numpy dynamic programme (trellis codes).  No text and no algorithm is shared between the two sides."""
import ctypes as C

import numpy as np
import p25_refdec as R
import p25_spec as S
import pytest
from test_oracle_fec import hostcheck  # noqa: F401  (fixture)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def corrupted_bch(rng, n):
    words = np.zeros(n, dtype=np.uint64)
    for i in range(n):
        w = S.bch_encode(int(rng.integers(0, 65536)))
        for p in rng.choice(63, int(rng.integers(0, 16)), replace=False):
            w ^= 1 << int(p)
        words[i] = w if i % 5 else int(rng.integers(0, 1 << 63))
    return words


def test_bch_against_exhaustive_search(oracle, hostcheck):
    O = oracle.lib()
    words = corrupted_bch(np.random.default_rng(11), 1500)
    data, nerr = R.bch_decode(words)
    assert (nerr >= 0).sum() > 400 and (nerr < 0).sum() > 400
    for i, w in enumerate(words):
        od, on, hd = C.c_uint16(), C.c_int(), C.c_uint32()
        ok = O.p25o_bch_decode(int(w), C.byref(od), C.byref(on))
        hn = hostcheck.hc_bch_decode(int(w), C.byref(hd))
        assert bool(ok) == (nerr[i] >= 0) and (hn >= 0) == (nerr[i] >= 0), i
        if ok:
            assert on.value == nerr[i] == hn and od.value == data[i] == hd.value, i


@pytest.mark.parametrize("name,bits", [("golay23", 23), ("golay24", 24), ("golay18", 18), ("hamming15", 15), ("hamming10", 10),
                                       ("cyclic16", 16)])
def test_short_codes_against_exhaustive_search(oracle, hostcheck, name, bits):
    O = oracle.lib()
    rng = np.random.default_rng(12)
    words = rng.integers(0, 1 << bits, 6000).astype(np.uint32)           # random words: every coset is exercised
    data, nerr = getattr(R, f"{name}_decode")(words)
    fo, fh = getattr(O, f"p25o_{name}_decode"), getattr(hostcheck, f"hc_{name}_decode")
    fh.argtypes = fo.argtypes
    for i, w in enumerate(words):
        a, b = C.c_uint32(), C.c_uint32()
        ra, rb = fo(int(w), C.byref(a)), fh(int(w), C.byref(b))
        assert ra == nerr[i] == rb, (name, i)
        if nerr[i] >= 0 or name != "golay24":      # a rejected Golay(24,12) word has no defined data
            assert a.value == data[i] == b.value, (name, i)


@pytest.mark.parametrize("n,k", [S.RS_SHORT, S.RS_MED, S.RS_LONG])
def test_reed_solomon_against_euclid(oracle, hostcheck, n, k):
    O = oracle.lib()
    t = (n - k) // 2
    rng = np.random.default_rng(13 + n + k)
    seen = set()
    for it in range(500):
        w = S.rs_encode([int(x) for x in rng.integers(0, 64, k)], n, k)
        for p in rng.choice(n, int(rng.integers(0, t + 3)), replace=False):
            w[int(p)] ^= int(rng.integers(1, 64))
        if it % 4 == 3:
            w = [int(x) for x in rng.integers(0, 64, n)]
        ref, rn = R.rs_decode(w, n, k)
        a = np.array(w, dtype=np.uint8)
        b = a.copy()
        ra, rb = O.p25o_rs_decode(_p(a), n, k), hostcheck.hc_rs_decode(_p(b), n, k)
        assert ra == rn == rb, (it, ra, rn, rb)
        assert list(a) == ref == list(b), it
        seen.add(rn)
    assert -1 in seen and t in seen and 0 in seen


def _trellis_blocks(rng, n, enc, nbytes, max_err):
    blocks = np.zeros((n, 98), dtype=np.uint8)
    for i in range(n):
        d = enc(rng.integers(0, 256, nbytes).astype(np.uint8).tobytes()).copy()
        for p in rng.choice(196, int(rng.integers(0, max_err)), replace=False):
            d[int(p) // 2] ^= 2 >> (int(p) & 1)
        blocks[i] = d if i % 5 else rng.integers(0, 4, 98)
    return blocks


def test_trellis_half_against_numpy_dp(oracle, hostcheck):
    O = oracle.lib()
    blocks = _trellis_blocks(np.random.default_rng(14), 1500, S.tsbk_block_dibits, 12, 20)
    ref, met = R.trellis_half_decode(blocks)
    assert (met >= 0).sum() > 300 and (met < 0).sum() > 250
    for i in range(len(blocks)):
        oa, ob = np.zeros(12, np.uint8), np.zeros(12, np.uint8)
        ra, rb = O.p25o_trellis_half_decode(_p(blocks[i]), _p(oa)), hostcheck.hc_trellis_half_decode(_p(blocks[i]), _p(ob))
        assert ra == met[i] == rb, i
        if ra >= 0:
            assert (oa == ref[i]).all() and (ob == ref[i]).all(), i


def test_trellis_34_round_trip_and_numpy_dp(oracle, hostcheck):
    O = oracle.lib()
    rng = np.random.default_rng(15)
    for _ in range(200):                                          # clean blocks decode to their payload, metric 0
        pl = rng.integers(0, 256, 18).astype(np.uint8)
        out = np.zeros(18, np.uint8)
        assert O.p25o_trellis_34_decode(_p(S.pdu_block34_dibits(pl.tobytes())), _p(out)) == 0 and (out == pl).all()
    blocks = _trellis_blocks(rng, 1500, S.pdu_block34_dibits, 18, 9)
    ref, met = R.trellis_34_decode(blocks)
    assert (met >= 0).sum() > 300 and (met[::5] < 0).all()       # random blocks never pass the metric bound
    for i in range(len(blocks)):
        oa, ob = np.zeros(18, np.uint8), np.zeros(18, np.uint8)
        ra, rb = O.p25o_trellis_34_decode(_p(blocks[i]), _p(oa)), hostcheck.hc_trellis_34_decode(_p(blocks[i]), _p(ob))
        assert ra == met[i] == rb, i
        if ra >= 0:
            assert (oa == ref[i]).all() and (ob == ref[i]).all(), i


def test_imbe_against_reference(oracle):
    O = oracle.lib()
    rng = np.random.default_rng(16)
    for it in range(300):
        d = S.imbe_encode([int(rng.integers(0, 1 << b)) for b in S.IMBE_U_BITS]).copy()
        for p in rng.choice(144, int(rng.integers(0, 14)), replace=False):
            d[int(p) // 2] ^= 2 >> (int(p) & 1)
        c, e = np.zeros(8, np.uint32), np.zeros(7, np.uint32)
        O.p25o_imbe_decode(_p(d), _p(c), _p(e))
        u, err = R.imbe_decode(d)
        assert list(c) == u and list(e) == err, it
