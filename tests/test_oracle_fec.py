"""Oracle FEC decoders: round trips against the independent numpy encoders of spec/p25_spec.py,
and a differential check of the CUDA library's decoder source (host build) against the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import p25_spec as S
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hostcheck():
    """tests/hostcheck: p25rx_b200/csrc/p25_fec.cuh compiled for the host -- test harness only."""
    d = os.path.join(HERE, "hostcheck")
    so = os.path.join(d, "libfec_hostcheck.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-w", "-o", so, os.path.join(d, "fec_hostcheck.cpp")])
    H = C.CDLL(so)
    H.hc_bch_decode.argtypes = [C.c_uint64, C.POINTER(C.c_uint32)]
    for n in ("golay23", "golay24", "golay18", "hamming15", "hamming10", "cyclic16"):
        getattr(H, f"hc_{n}_decode").argtypes = [C.c_uint32, C.POINTER(C.c_uint32)]
    return H


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_bch_round_trip_and_radius(oracle, hostcheck):
    O = oracle.lib()
    rng = np.random.default_rng(0)
    for it in range(1500):
        d = int(rng.integers(0, 65536))
        ne = int(rng.integers(0, 16))
        w = S.bch_encode(d)
        for p in rng.choice(63, ne, replace=False):
            w ^= 1 << int(p)
        if it % 7 == 6:
            w = int(rng.integers(0, 1 << 63))
        od, on = C.c_uint16(), C.c_int()
        ok = O.p25o_bch_decode(w, C.byref(od), C.byref(on))
        if ne <= 11 and it % 7 != 6:
            assert ok and od.value == d and on.value == ne
        hd = C.c_uint32()
        hn = hostcheck.hc_bch_decode(w, C.byref(hd))
        assert (hn >= 0) == bool(ok)
        if ok:
            assert hn == on.value and hd.value == od.value


@pytest.mark.parametrize("name,bits,dbits,t", [("golay23", 23, 12, 3), ("golay24", 24, 12, 3), ("golay18", 18, 6, 3),
                                               ("hamming15", 15, 11, 1), ("hamming10", 10, 6, 1), ("cyclic16", 16, 8, 2)])
def test_short_codes(oracle, hostcheck, name, bits, dbits, t):
    O = oracle.lib()
    enc = getattr(S, f"{name}_encode")
    fo, fh = getattr(O, f"p25o_{name}_decode"), getattr(hostcheck, f"hc_{name}_decode")
    rng = np.random.default_rng(3)
    for it in range(3000):
        d = int(rng.integers(0, 1 << dbits))
        ne = int(rng.integers(0, t + 3))
        w = enc(d)
        for p in rng.choice(bits, ne, replace=False):
            w ^= 1 << int(p)
        a, b = C.c_uint32(), C.c_uint32()
        ra, rb = fo(w, C.byref(a)), fh(w, C.byref(b))
        if ne <= t:
            assert ra == ne and a.value == d, (name, ne, ra)
        assert ra == rb and a.value == b.value
    if bits <= 16:   # exhaustive differential
        for w in range(1 << bits):
            a, b = C.c_uint32(), C.c_uint32()
            assert fo(w, C.byref(a)) == fh(w, C.byref(b)) and a.value == b.value


@pytest.mark.parametrize("n,k", [S.RS_SHORT, S.RS_MED, S.RS_LONG])
def test_reed_solomon(oracle, hostcheck, n, k):
    O = oracle.lib()
    t = (n - k) // 2
    rng = np.random.default_rng(n * 100 + k)
    for it in range(1200):
        cw = S.rs_encode([int(x) for x in rng.integers(0, 64, k)], n, k)
        ne = int(rng.integers(0, t + 3))
        w = list(cw)
        for p in rng.choice(n, ne, replace=False):
            w[int(p)] ^= int(rng.integers(1, 64))
        if it % 4 == 3:
            w = [int(x) for x in rng.integers(0, 64, n)]
        a = np.array(w, dtype=np.uint8)
        b = a.copy()
        ra, rb = O.p25o_rs_decode(_p(a), n, k), hostcheck.hc_rs_decode(_p(b), n, k)
        if ne <= t and it % 4 != 3:
            assert ra == ne and list(a) == cw
        assert ra == rb and (a == b).all()
        if ra < 0:
            assert list(a) == w      # an unrecoverable word is returned untouched


def test_trellis(oracle, hostcheck):
    O = oracle.lib()
    rng = np.random.default_rng(5)
    for it in range(2000):
        pl = rng.integers(0, 256, 12).astype(np.uint8).tobytes()
        d = S.tsbk_block_dibits(pl).copy()
        ne = int(rng.integers(0, 14))
        for p in rng.choice(196, ne, replace=False):
            d[int(p) // 2] ^= 2 >> (int(p) & 1)
        if it % 5 == 4:
            d = rng.integers(0, 4, 98).astype(np.uint8)
        oa, ob = np.zeros(12, np.uint8), np.zeros(12, np.uint8)
        ra, rb = O.p25o_trellis_half_decode(_p(d), _p(oa)), hostcheck.hc_trellis_half_decode(_p(d), _p(ob))
        if ne <= 2 and it % 5 != 4:   # free distance 5 bits -> two errors always corrected
            assert ra == ne and oa.tobytes() == pl
        if it % 5 == 4:
            assert ra == -1               # random blocks are rejected by the metric bound
        assert ra == rb and (ra < 0 or (oa == ob).all())


def test_imbe(oracle, hostcheck):
    O = oracle.lib()
    rng = np.random.default_rng(6)
    for it in range(1500):
        u = [int(rng.integers(0, 1 << b)) for b in S.IMBE_U_BITS]
        d = S.imbe_encode(u).copy()
        ne = int(rng.integers(0, 10))
        for p in rng.choice(144, ne, replace=False):
            d[int(p) // 2] ^= 2 >> (int(p) & 1)
        ca, ea = np.zeros(8, np.uint32), np.zeros(7, np.uint32)
        cb, eb = ca.copy(), ea.copy()
        O.p25o_imbe_decode(_p(d), _p(ca), _p(ea))
        hostcheck.hc_imbe_decode(_p(d), _p(cb), _p(eb))
        if ne == 0:
            assert list(ca) == u and ea.sum() == 0
        assert (ca == cb).all() and (ea == eb).all()


def test_crc(oracle):
    O = oracle.lib()
    rng = np.random.default_rng(7)
    for _ in range(100):
        m = rng.integers(0, 256, 10).astype(np.uint8)
        assert O.p25o_crc_ccitt(_p(m), 10) == S.crc_ccitt_p25(m.tobytes())
