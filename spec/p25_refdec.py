"""Reference decoders for the P25 codes, written to share NO text and NO algorithm with the oracle's C++ decoders
(oracle/p25_oracle.cpp) or the device decoders (p25rx_b200/csrc/p25_fec.cuh).  TEST INFRASTRUCTURE ONLY.

The C++ oracle and the CUDA kernels both decode algebraically (syndromes, Berlekamp-Massey, Chien, Forney, syndrome
tables); comparing them with each other cannot expose a shared misreading.  Everything here works from the ENCODERS
of spec/p25_spec.py instead:

  * binary block codes (BCH(63,16,23), Golay(23,12), Golay(24,12), Golay(18,6), Hamming(15,11), Hamming(10,6),
    cyclic(16,8)): exhaustive nearest-code-word search over the full code book, accepted within the bounded
    distance the receiver corrects (11 / 3 / 3 / 3 / 1 / 1 / 2).  Bounded-distance decoding has a unique answer,
    so any correct decoder must agree bit for bit, including on which words are rejected;
  * Reed-Solomon over GF(64): Sugiyama's Euclidean algorithm for the key equation and a Vandermonde solve for the
    error values (the oracle and the kernels use Berlekamp-Massey and Forney);
  * trellis codes: a vectorised dynamic programme over whole batches (numpy argmin keeps the lowest predecessor on ties,
    the rule the receiver documents);
  * IMBE frames: de-interleave by the schedule, PN masks from the generator, the block decoders above.
"""
from __future__ import annotations

import numpy as np

import p25_spec as S


def _popcount(a: np.ndarray) -> np.ndarray:
    return np.bitwise_count(a)


# --------------------------------------------------------------------------- binary block codes: full code books
_BOOKS: dict[str, np.ndarray] = {}


def _book(name: str) -> np.ndarray:
    if name not in _BOOKS:
        enc, kbits = {"bch": (S.bch_encode, 16), "golay23": (S.golay23_encode, 12), "golay24": (S.golay24_encode, 12),
                      "hamming15": (S.hamming15_encode, 11), "hamming10": (S.hamming10_encode, 6),
                      "cyclic16": (S.cyclic16_encode, 8)}[name]
        _BOOKS[name] = np.array([enc(d) for d in range(1 << kbits)], dtype=np.uint64)
    return _BOOKS[name]


def _nearest(name: str, words: np.ndarray, chunk: int = 64):
    """(index of the nearest code word, its distance) for every word; ties cannot occur inside the decoding radius."""
    book = _book(name)
    words = np.asarray(words, dtype=np.uint64)
    idx = np.zeros(len(words), dtype=np.int64)
    dist = np.zeros(len(words), dtype=np.int64)
    for i in range(0, len(words), chunk):
        d = _popcount(book[None, :] ^ words[i:i + chunk, None])
        idx[i:i + chunk] = np.argmin(d, axis=1)
        dist[i:i + chunk] = d[np.arange(d.shape[0]), idx[i:i + chunk]]
    return idx, dist


def bch_decode(words63: np.ndarray):
    """-> (data16, nerr) with nerr = -1 where no code word lies within 11 bits."""
    idx, dist = _nearest("bch", np.asarray(words63, dtype=np.uint64) & np.uint64((1 << 63) - 1))
    ok = dist <= S.BCH_T
    return np.where(ok, idx, 0).astype(np.uint32), np.where(ok, dist, -1).astype(np.int32)


def golay23_decode(words: np.ndarray):
    """Perfect code: every word is within 3 bits of exactly one code word."""
    idx, dist = _nearest("golay23", np.asarray(words, dtype=np.uint64) & np.uint64(0x7FFFFF))
    assert (dist <= 3).all()
    return idx.astype(np.uint32), dist.astype(np.int32)


def golay24_decode(words: np.ndarray):
    """Extended code, distance 8: up to 3 bits corrected, anything farther rejected.  A rejected word still reports the
    data bits of the 23-bit decode (the receiver passes them on); callers compare data only where nerr >= 0 unless they
    model that too, which golay24_raw_data does."""
    w = np.asarray(words, dtype=np.uint64) & np.uint64(0xFFFFFF)
    idx, dist = _nearest("golay24", w)
    ok = dist <= 3
    return idx.astype(np.uint32), np.where(ok, dist, -1).astype(np.int32)


def golay18_decode(words: np.ndarray):
    """Shortened (18,6): the six leading data bits of the (24,12) code word are zero.  Rejected words report the raw
    data field (bits 17..12)."""
    w = np.asarray(words, dtype=np.uint64) & np.uint64(0x3FFFF)
    idx, dist = _nearest("golay24", w)
    ok = (dist <= 3) & ((idx >> 6) == 0)
    data = np.where(ok, idx & 0x3F, (w >> np.uint64(12)).astype(np.int64) & 0x3F)
    return data.astype(np.uint32), np.where(ok, dist, -1).astype(np.int32)


def hamming15_decode(words: np.ndarray):
    """Perfect single-error-correcting code."""
    idx, dist = _nearest("hamming15", np.asarray(words, dtype=np.uint64) & np.uint64(0x7FFF))
    assert (dist <= 1).all()
    return idx.astype(np.uint32), dist.astype(np.int32)


def hamming10_decode(words: np.ndarray):
    """Shortened (10,6,3): words farther than one bit from every code word are rejected and report the raw data bits."""
    w = np.asarray(words, dtype=np.uint64) & np.uint64(0x3FF)
    idx, dist = _nearest("hamming10", w)
    ok = dist <= 1
    data = np.where(ok, idx, (w >> np.uint64(4)).astype(np.int64))
    return data.astype(np.uint32), np.where(ok, dist, -1).astype(np.int32)


def cyclic16_decode(words: np.ndarray):
    """(16,8,5): up to two bits corrected; rejected words report the raw data byte."""
    w = np.asarray(words, dtype=np.uint64) & np.uint64(0xFFFF)
    idx, dist = _nearest("cyclic16", w)
    ok = dist <= 2
    data = np.where(ok, idx, (w >> np.uint64(8)).astype(np.int64))
    return data.astype(np.uint32), np.where(ok, dist, -1).astype(np.int32)


# --------------------------------------------------------------------------- Reed-Solomon: Euclid + Vandermonde
def _gf_mul(a: int, b: int) -> int:
    return S.gf_mul(a, b)


def _gf_inv(a: int) -> int:
    return S.gf_inv(a)


def _poly_trim(p: list[int]) -> list[int]:
    while len(p) > 1 and p[-1] == 0:
        p = p[:-1]
    return p


def _poly_divmod(a: list[int], b: list[int]):
    """a = q b + r over GF(64); index = degree."""
    a = list(a)
    b = _poly_trim(list(b))
    db = len(b) - 1
    q = [0] * max(1, len(a) - db)
    inv = _gf_inv(b[-1])
    for d in range(len(a) - 1, db - 1, -1):
        c = _gf_mul(a[d], inv)
        if c:
            q[d - db] = c
            for i, bc in enumerate(b):
                a[d - db + i] ^= _gf_mul(c, bc)
    return _poly_trim(q), _poly_trim(a[:db] if db else [0])


def _poly_mul(a: list[int], b: list[int]) -> list[int]:
    out = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                out[i + j] ^= _gf_mul(x, y)
    return out


def _poly_add(a: list[int], b: list[int]) -> list[int]:
    n = max(len(a), len(b))
    return [(a[i] if i < len(a) else 0) ^ (b[i] if i < len(b) else 0) for i in range(n)]


def _poly_eval(p: list[int], x: int) -> int:
    acc = 0
    for c in reversed(p):
        acc = _gf_mul(acc, x) ^ c
    return acc


def rs_decode(sym: list[int], n: int, k: int):
    """-> (corrected symbols, number of corrected symbols) or (the word as received, -1).  sym[0] = highest degree."""
    nroots = n - k
    t = nroots // 2
    alpha = lambda e: int(S.GF_EXP[e % 63])
    r_poly = list(reversed([int(x) for x in sym]))              # index = degree
    synd = [_poly_eval(r_poly, alpha(j)) for j in range(1, nroots + 1)]
    if not any(synd):
        return list(sym), 0
    # Sugiyama: run Euclid on x^(2t) and S(x) until the remainder's degree drops below t
    r0, r1 = [0] * nroots + [1], _poly_trim(list(synd))
    t0, t1 = [0], [1]
    while len(_poly_trim(r1)) - 1 >= t and any(r1):
        q, rem = _poly_divmod(r0, r1)
        r0, r1 = r1, rem
        t0, t1 = t1, _poly_trim(_poly_add(t0, _poly_mul(q, t1)))
    lam = _poly_trim(t1)
    if lam[0] == 0:
        return list(sym), -1
    nu = len(lam) - 1
    if nu == 0 or nu > t:
        return list(sym), -1
    pos = [p for p in range(63) if _poly_eval(lam, alpha(63 - p)) == 0]   # error at x^p  <=>  lambda(alpha^-p) = 0
    if len(pos) != nu or any(p >= n for p in pos):
        return list(sym), -1
    # error values from the first nu syndrome equations: sum_i e_i alpha^(j p_i) = S_j  (Gaussian elimination)
    A = [[alpha(j * p) for p in pos] + [synd[j - 1]] for j in range(1, nu + 1)]
    for c in range(nu):
        piv = next((r for r in range(c, nu) if A[r][c]), None)
        if piv is None:
            return list(sym), -1
        A[c], A[piv] = A[piv], A[c]
        inv = _gf_inv(A[c][c])
        A[c] = [_gf_mul(x, inv) for x in A[c]]
        for r in range(nu):
            if r != c and A[r][c]:
                f = A[r][c]
                A[r] = [x ^ _gf_mul(f, y) for x, y in zip(A[r], A[c])]
    vals = [A[i][nu] for i in range(nu)]
    if any(v == 0 for v in vals):
        return list(sym), -1
    out = list(sym)
    for p, v in zip(pos, vals):
        out[n - 1 - p] ^= v
    chk = list(reversed(out))
    if any(_poly_eval(chk, alpha(j)) for j in range(1, nroots + 1)):
        return list(sym), -1
    return out, nu


# --------------------------------------------------------------------------- trellis codes: batched dynamic programme
def _deinterleave(blocks: np.ndarray) -> np.ndarray:
    """[B][98] received dibits -> [B][49] 4-bit symbols in trellis order."""
    b = np.asarray(blocks, dtype=np.int64).reshape(-1, 98)
    slots = (b[:, 0::2] << 2) | b[:, 1::2]
    return slots[:, np.array(S.interleave_perm())]


def _viterbi(sym: np.ndarray, expect: np.ndarray, bound: int):
    """sym [B][49]; expect[ps][ns] = 4-bit pair emitted on the transition ps -> ns.  Returns (inputs [B][48], metric or -1)."""
    B, ns = sym.shape[0], expect.shape[0]
    INF = 1 << 20
    m = np.full((B, ns), INF, dtype=np.int64)
    m[:, 0] = 0
    frm = np.zeros((49, B, ns), dtype=np.int64)
    pc = np.array([bin(i).count("1") for i in range(16)], dtype=np.int64)
    for i in range(49):
        cost = m[:, :, None] + pc[expect[None, :, :] ^ sym[:, i, None, None]]      # [B][ps][ns]
        frm[i] = np.argmin(cost, axis=1)                                            # first minimum = lowest predecessor
        m = np.min(cost, axis=1)
    metric = m[:, 0].copy()
    st = np.zeros(B, dtype=np.int64)
    inputs = np.zeros((B, 49), dtype=np.int64)
    for i in range(48, -1, -1):
        inputs[:, i] = st
        st = frm[i][np.arange(B), st]
    return inputs[:, :48], np.where(metric <= bound, metric, -1).astype(np.int32)


def trellis_half_decode(blocks: np.ndarray):
    """[B][98] dibits -> ([B][12] bytes, metric or -1)."""
    expect = np.array(S.CONSTELLATION)[np.array(S.TRELLIS_HALF)]      # [state][input]; next state = input
    inp, met = _viterbi(_deinterleave(blocks), expect, S.VITERBI_MAX_FIX)
    bits = np.stack([(inp >> 1) & 1, inp & 1], axis=2).reshape(len(inp), 96).astype(np.uint8)
    return np.packbits(bits, axis=1), met


def trellis_34_decode(blocks: np.ndarray):
    """[B][98] dibits -> ([B][18] bytes, metric or -1)."""
    expect = np.array(S.CONSTELLATION)[np.array(S.TRELLIS_3_4)]
    inp, met = _viterbi(_deinterleave(blocks), expect, S.VITERBI34_MAX_FIX)
    bits = np.stack([(inp >> 2) & 1, (inp >> 1) & 1, inp & 1], axis=2).reshape(len(inp), 144).astype(np.uint8)
    return np.packbits(bits, axis=1), met


# --------------------------------------------------------------------------- IMBE frame
def imbe_decode(dibits72: np.ndarray):
    """-> (u0..u7, 7 corrected-bit counts)."""
    d = np.asarray(dibits72, dtype=np.int64)
    bits = np.stack([(d >> 1) & 1, d & 1], axis=1).reshape(-1)
    cw = [0] * 8
    for (c, b), bit in zip(S.imbe_schedule(), bits):
        cw[int(c)] |= int(bit) << int(b)
    u, err = [0] * 8, [0] * 7
    dd, ee = golay23_decode(np.array([cw[0]]))
    u[0], err[0] = int(dd[0]), int(ee[0])
    masks = S.imbe_pn_masks(u[0])
    for c in range(1, 7):
        w = np.array([cw[c] ^ masks[c]])
        dd, ee = golay23_decode(w) if c < 4 else hamming15_decode(w)
        u[c], err[c] = int(dd[0]), int(ee[0])
    u[7] = cw[7] & 0x7F
    return u, err
