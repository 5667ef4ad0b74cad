"""Algebraic self-tests of the [STD] constants in spec/p25_spec.py.

No copy of TIA-102.BAAA-A and no reference vector is available, so every
recalled constant is pinned by a property that a wrong value would violate."""
import itertools

import numpy as np
import p25_spec as S


def test_bch_generator_is_the_t11_bch_code():
    # the recalled octal generator equals lcm(minpoly(alpha^1..alpha^22)) over GF(64), x^6+x+1
    assert S.bch_derived_generator() == S.BCH_GEN
    assert S.BCH_GEN.bit_length() - 1 == 47
    assert S._polymod2((1 << 63) | 1, S.BCH_GEN) == 0          # divides x^63 + 1


def test_bch_min_weight_of_basis_and_pairs():
    rows = [S.bch_encode(1 << i) for i in range(16)]
    w = min(bin(r).count("1") for r in rows)
    for a, b in itertools.combinations(rows, 2):
        w = min(w, bin(a ^ b).count("1"))
    assert w >= 23


def test_golay_generator_and_distance():
    assert S._polymod2((1 << 23) | 1, S.GOLAY_GEN) == 0
    weights = [bin(S.golay23_encode(d)).count("1") for d in range(1, 4096)]
    assert min(weights) == 7
    assert min(bin(S.golay24_encode(d)).count("1") for d in range(1, 4096)) == 8
    assert min(bin(S.golay18_encode(d)).count("1") for d in range(1, 64)) == 8
    S.golay23_syndrome_table()   # asserts the code is perfect (every syndrome has one coset leader)


def test_hamming_distances():
    assert min(bin(S.hamming15_encode(d)).count("1") for d in range(1, 2048)) == 3
    assert min(bin(S.hamming10_encode(d)).count("1") for d in range(1, 64)) == 3
    assert sorted(S.HAMMING15_COLS + [1, 2, 4, 8]) == list(range(1, 16))


def test_cyclic_distance():
    assert S._polymod2((1 << 17) | 1, S.CYCLIC_GEN) == 0
    assert min(bin(S.cyclic16_encode(d)).count("1") for d in range(1, 256)) == 5


def test_rs_codewords_have_zero_syndromes():
    rng = np.random.default_rng(0)
    for n, k in (S.RS_SHORT, S.RS_MED, S.RS_LONG):
        cw = S.rs_encode([int(x) for x in rng.integers(0, 64, k)], n, k)
        for j in range(1, n - k + 1):
            acc = 0
            for c in cw:
                acc = S.gf_mul(acc, int(S.GF_EXP[j])) ^ c
            assert acc == 0
        # minimum distance n-k+1: a single data symbol produces a full-weight parity section
        w = min(sum(1 for c in S.rs_encode([v if i == p else 0 for i in range(k)], n, k) if c)
                for p in range(k) for v in (1, 17, 63))
        assert w >= n - k + 1


def test_trellis_tables():
    for tab, n_in in ((S.TRELLIS_HALF, 4), (S.TRELLIS_3_4, 8)):
        for row in tab:
            assert len(set(row)) == n_in                      # distinct outputs per state
    assert sorted(p for row in S.TRELLIS_HALF for p in row) == list(range(16))
    assert sorted(S.CONSTELLATION) == list(range(16))
    # free distance (in bits) of the half-rate code from the all-zero path, depth-limited search
    best = 99
    for first in (1, 2, 3):
        frontier = {first: bin(S.CONSTELLATION[S.TRELLIS_HALF[0][first]] ^ S.CONSTELLATION[S.TRELLIS_HALF[0][0]]).count("1")}
        for _ in range(6):
            nxt = {}
            for st, d in frontier.items():
                for inp in range(4):
                    dd = d + bin(S.CONSTELLATION[S.TRELLIS_HALF[st][inp]] ^ S.CONSTELLATION[S.TRELLIS_HALF[0][0]]).count("1")
                    if inp == 0:
                        best = min(best, dd)
                    elif dd < nxt.get(inp, 99):
                        nxt[inp] = dd
            frontier = nxt
    assert best == 5


def test_interleaver_and_layouts():
    perm = S.interleave_perm()
    assert sorted(perm) == list(range(49)) and perm[:8] == [0, 13, 25, 37, 1, 14, 26, 38] and perm[-1] == 12
    sched = S.imbe_schedule()
    assert len({(int(c), int(b)) for c, b in sched}) == 144
    assert [tuple(x) for x in sched[:4].tolist()] == [(0, 22), (1, 22), (2, 22), (3, 22)]
    assert S.LDU_DIBITS == 784 and 24 + 32 + S.LDU_DIBITS == 24 * 35
    assert (24 + 32 + S.HDU_DIBITS + 5) % 35 == 0 and (24 + 32 + S.TDULC_DIBITS + 10) % 35 == 0
    assert (24 + 32 + 3 * S.TSBK_DIBITS) % 35 == 0


def test_frame_sync_and_fingerprint():
    assert S.FRAME_SYNC == 0x5575F5FF77FF
    syms = S.frame_sync_symbols()
    assert set(syms.tolist()) == {3, -3} and (syms > 0).sum() == 11
    fp = S.sync_fingerprint()
    assert len(fp) == 231 and np.allclose(fp[::10], syms / 3.0, atol=1e-6)


def test_crc_residue():
    rng = np.random.default_rng(1)
    msg = rng.integers(0, 256, 10).astype(np.uint8).tobytes()
    crc = S.crc_ccitt_p25(msg)
    # linearity: the (un-inverted) remainder of message || remainder is zero
    rem = crc ^ 0xFFFF
    assert S.crc_ccitt_p25(msg + bytes([rem >> 8, rem & 0xFF])) ^ 0xFFFF == 0


def test_generated_headers_are_current():
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert subprocess.call([sys.executable, os.path.join(root, "spec", "gen_tables.py"), "--check"]) == 0
