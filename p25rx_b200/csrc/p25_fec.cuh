// p25_fec.cuh -- P25 Phase 1 forward-error-correction decoders as per-thread device functions.
//
// These replace the decoders the reference reaches through p25::MessageReceiver::feed
// (reference src/recv.rs:207, src/replay.rs:44; crates p25 @a96c564, cai_golay 0.1.1,
// cai_cyclic 0.1.2, binfield_matrix 0.2.0 -- not vendored).  Algorithms are
// the published CAI ones (TIA-102.BAAA-A); all tables come from p25_tables.h, generated from
// spec/p25_spec.py.  Every function is a plain function of its arguments so that one thread
// can decode one code word; the walker kernel calls them from the lane that owns a word.
//
// P25_FN expands to __device__ for the product build.  tests/hostcheck compiles this same
// header for the host (P25_FEC_HOSTCHECK) purely to unit-test the decoder logic on machines
// without a GPU; that build is never loaded by the p25rx_b200 package.
#pragma once
#include <stdint.h>

#include "p25_tables.h"

#ifdef P25_FEC_HOSTCHECK
#define P25_FN static inline
#define P25_POPC(x) __builtin_popcount(x)
#define P25_ROLLED
#else
#define P25_FN __device__ __forceinline__
#define P25_POPC(x) __popc(x)
// The walker is instruction-fetch bound on voice traffic (stall_no_inst 42 %: ~150 KB of code, every warp in a different
// decoder); the short bit loops below stay rolled so that the decoders fit the instruction cache.
#define P25_ROLLED _Pragma("unroll 1")
#endif

// Device-resident copy of the decode tables (filled from p25_tables.h at context creation;
// the walker kernel stages it into shared memory).
struct alignas(16) P25DevTables {
    const uint32_t* golay_syn;   // 2048 entries; stays in global memory (8 KB, L1/L2 resident, voice paths only) so that
    uint64_t pad_;               // the shared-memory copy of this struct is ~2 KB and walker CTAs co-reside with ddc_fm
    uint16_t cyclic_syn[256];
    uint16_t ham15_syn[16];
    uint16_t ham10_syn[16];
    uint16_t ldu_kind[16], ldu_start[16], ldu_len[16];
    uint8_t gf_exp[128];
    uint8_t gf_log[64];
    uint8_t ham15_cols[16];
    uint8_t ham10_cols[8];
    uint8_t trellis_pair[16];   // [prev state * 4 + next state] -> expected 4-bit dibit pair
    uint8_t trellis34_pair[64]; // 3/4-rate: [prev state * 8 + next state] -> expected 4-bit dibit pair
    uint8_t interleave[52];
    uint8_t imbe_cw[144];
    uint8_t imbe_bit[144];
    uint8_t imbe_cw_bits[8];
    uint64_t bch_par[22][6];    // BCH(63,16,23) syndromes as parities: bit b of S_j = parity(word & bch_par[j - 1][b])
};

static inline void p25_fill_tables(P25DevTables* t) {
    t->golay_syn = P25_GOLAY23_SYN;   // host table; the CUDA library repoints this at its device copy
    t->pad_ = 0;
    for (int i = 0; i < 256; i++) t->cyclic_syn[i] = P25_CYCLIC16_SYN[i];
    for (int i = 0; i < 16; i++) {
        t->ham15_syn[i] = P25_HAMMING15_SYN[i];
        t->ham10_syn[i] = P25_HAMMING10_SYN[i];
        t->ldu_kind[i] = P25_LDU_PART_KIND[i];
        t->ldu_start[i] = P25_LDU_PART_START[i];
        t->ldu_len[i] = P25_LDU_PART_LEN[i];
        t->ham15_cols[i] = i < 11 ? P25_HAMMING15_COLS[i] : 0;
        t->trellis_pair[i] = P25_CONSTELLATION[P25_TRELLIS_HALF[i]];
    }
    for (int i = 0; i < 64; i++) t->trellis34_pair[i] = P25_CONSTELLATION[P25_TRELLIS_3_4[i]];
    for (int j = 1; j <= 22; j++)           // S_j = sum_i r_i alpha^(j i): bit b of alpha^(j i) selects r_i into parity b
        for (int b = 0; b < 6; b++) {
            uint64_t m = 0;
            for (int i = 0; i < 63; i++)
                if ((P25_GF_EXP[(j * i) % 63] >> b) & 1) m |= (uint64_t)1 << i;
            t->bch_par[j - 1][b] = m;
        }
    for (int i = 0; i < 128; i++) t->gf_exp[i] = P25_GF_EXP[i];
    for (int i = 0; i < 64; i++) t->gf_log[i] = P25_GF_LOG[i];
    for (int i = 0; i < 8; i++) {
        t->ham10_cols[i] = i < 6 ? P25_HAMMING10_COLS[i] : 0;
        t->imbe_cw_bits[i] = P25_IMBE_CW_BITS[i];
    }
    for (int i = 0; i < 52; i++) t->interleave[i] = i < 49 ? P25_INTERLEAVE[i] : 0;
    for (int i = 0; i < 144; i++) {
        t->imbe_cw[i] = P25_IMBE_SCHED_CW[i];
        t->imbe_bit[i] = P25_IMBE_SCHED_BIT[i];
    }
}

// ------------------------------------------------------------------ GF(2^6)
P25_FN int p25_gf_mul(const P25DevTables& T, int a, int b) {
    return (a && b) ? T.gf_exp[T.gf_log[a] + T.gf_log[b]] : 0;
}
P25_FN int p25_gf_div(const P25DevTables& T, int a, int b) {  // b != 0
    return a ? T.gf_exp[T.gf_log[a] + 63 - T.gf_log[b]] : 0;
}
P25_FN int p25_gf_alpha(const P25DevTables& T, int e) {  // alpha^e, 0 <= e < 126
    return T.gf_exp[e];
}

// Berlekamp-Massey: S[0..n-1] = S_1..S_n; lam[0..n] receives the locator.  Returns its length L.
template <int N>
P25_FN int p25_berlekamp_massey(const P25DevTables& T, const uint8_t* S, uint8_t* lam) {
    uint8_t B[N + 1], Tm[N + 1];
    for (int i = 0; i <= N; i++) {
        lam[i] = 0;
        B[i] = 0;
    }
    lam[0] = 1;
    B[0] = 1;
    int L = 0, m = 1, b = 1;
    for (int r = 0; r < N; r++) {
        int d = S[r];
        for (int i = 1; i <= L; i++) d ^= p25_gf_mul(T, lam[i], S[r - i]);
        if (d == 0) {
            m++;
            continue;
        }
        const int coef = p25_gf_div(T, d, b);
        if (2 * L <= r) {
            for (int i = 0; i <= N; i++) Tm[i] = lam[i];
            for (int i = 0; i + m <= N; i++) lam[i + m] ^= (uint8_t)p25_gf_mul(T, coef, B[i]);
            L = r + 1 - L;
            for (int i = 0; i <= N; i++) B[i] = Tm[i];
            b = d;
            m = 1;
        } else {
            for (int i = 0; i + m <= N; i++) lam[i + m] ^= (uint8_t)p25_gf_mul(T, coef, B[i]);
            m++;
        }
    }
    return L;
}

P25_FN int p25_poly_eval(const P25DevTables& T, const uint8_t* p, int deg, int x) {
    int acc = 0;
    P25_ROLLED
    for (int i = deg; i >= 0; i--) acc = p25_gf_mul(T, acc, x) ^ p[i];
    return acc;
}

// ------------------------------------------------------------------ BCH(63,16,23)
P25_FN bool p25_bch_syndromes(const P25DevTables& T, uint64_t w, uint8_t* S) {
    // odd syndromes by summation, even ones by squaring (S_2j = S_j^2 for a binary word)
    int any = 0;
    for (int j = 1; j <= 2 * P25_BCH_T; j += 2) {
        int acc = 0, e = 0;  // e = (j * i) mod 63
        for (int i = 0; i < 63; i++) {
            if ((w >> i) & 1) acc ^= T.gf_exp[e];
            e += j;
            if (e >= 63) e -= 63;
        }
        S[j - 1] = (uint8_t)acc;
        any |= acc;
    }
    for (int j = 2; j <= 2 * P25_BCH_T; j += 2) {
        const int h = S[j / 2 - 1];
        S[j - 1] = (uint8_t)p25_gf_mul(T, h, h);
    }
    return any == 0;
}

// word63: bit i = coefficient of x^i.  Returns corrected-bit count (0..11) or -1.
P25_FN int p25_bch_decode(const P25DevTables& T, uint64_t word63, uint32_t* data16) {
    uint64_t w = word63 & 0x7FFFFFFFFFFFFFFFULL;
    uint8_t S[2 * P25_BCH_T];
    int fixed = 0;
    if (!p25_bch_syndromes(T, w, S)) {
        uint8_t lam[2 * P25_BCH_T + 1];
        const int L = p25_berlekamp_massey<2 * P25_BCH_T>(T, S, lam);
        if (L > P25_BCH_T) return -1;
        int roots = 0;
        uint64_t flip = 0;
        for (int p = 0; p < 63; p++) {
            if (p25_poly_eval(T, lam, L, p25_gf_alpha(T, (63 - p) % 63)) == 0) {
                flip |= 1ULL << p;
                roots++;
            }
        }
        if (roots != L) return -1;
        w ^= flip;
        if (!p25_bch_syndromes(T, w, S)) return -1;
        fixed = L;
    }
    *data16 = (uint32_t)(w >> 47);
    return fixed;
}

// ------------------------------------------------------------------ Golay / Hamming / cyclic
P25_FN uint32_t p25_polymod(uint32_t a, uint32_t g, int gdeg, int abits) {
    P25_ROLLED
    for (int i = abits - 1; i >= gdeg; i--)
        if ((a >> i) & 1) a ^= g << (i - gdeg);
    return a;
}

P25_FN int p25_golay23_decode(const P25DevTables& T, uint32_t word, uint32_t* data12) {
    word &= 0x7FFFFF;
    const uint32_t e = T.golay_syn[p25_polymod(word, P25_GOLAY_GEN, 11, 23)];
    *data12 = (word ^ e) >> 11;
    return P25_POPC(e);
}

P25_FN int p25_golay24_decode(const P25DevTables& T, uint32_t word, uint32_t* data12) {
    word &= 0xFFFFFF;
    const uint32_t w23 = word >> 1;
    const uint32_t e = T.golay_syn[p25_polymod(w23, P25_GOLAY_GEN, 11, 23)];
    const uint32_t c23 = w23 ^ e;
    int nerr = P25_POPC(e);
    if ((uint32_t)(P25_POPC(c23) & 1) != (word & 1)) nerr++;
    *data12 = c23 >> 11;
    return nerr > 3 ? -1 : nerr;
}

P25_FN int p25_golay18_decode(const P25DevTables& T, uint32_t word, uint32_t* data6) {
    uint32_t d12;
    const int n = p25_golay24_decode(T, word & 0x3FFFF, &d12);
    if (n < 0 || (d12 >> 6)) {
        *data6 = (word >> 12) & 0x3F;
        return -1;
    }
    *data6 = d12 & 0x3F;
    return n;
}

P25_FN int p25_hamming15_decode(const P25DevTables& T, uint32_t word, uint32_t* data11) {
    word &= 0x7FFF;
    uint32_t p = 0;
    P25_ROLLED
    for (int i = 0; i < 11; i++)
        if ((word >> (14 - i)) & 1) p ^= T.ham15_cols[i];
    const uint32_t s = p ^ (word & 0xF);
    word ^= T.ham15_syn[s];
    *data11 = word >> 4;
    return s ? 1 : 0;
}

P25_FN int p25_hamming10_decode(const P25DevTables& T, uint32_t word, uint32_t* data6) {
    word &= 0x3FF;
    uint32_t p = 0;
    P25_ROLLED
    for (int i = 0; i < 6; i++)
        if ((word >> (9 - i)) & 1) p ^= T.ham10_cols[i];
    const uint32_t s = p ^ (word & 0xF);
    const uint32_t flip = T.ham10_syn[s];
    if (flip == 0xFFFF) {
        *data6 = word >> 4;
        return -1;
    }
    *data6 = (word ^ flip) >> 4;
    return s ? 1 : 0;
}

P25_FN int p25_cyclic16_decode(const P25DevTables& T, uint32_t word, uint32_t* data8) {
    word &= 0xFFFF;
    const uint32_t e = T.cyclic_syn[p25_polymod(word, P25_CYCLIC_GEN, 8, 16)];
    if (e == 0xFFFF) {
        *data8 = word >> 8;
        return -1;
    }
    *data8 = (word ^ e) >> 8;
    return P25_POPC(e);
}

// ------------------------------------------------------------------ Reed-Solomon over GF(64)
P25_FN bool p25_rs_syndromes(const P25DevTables& T, const uint8_t* sym, int n, int nroots, uint8_t* S) {
    int any = 0;
    for (int j = 1; j <= nroots; j++) {
        const int a = T.gf_exp[j];
        int acc = 0;
        for (int i = 0; i < n; i++) acc = p25_gf_mul(T, acc, a) ^ sym[i];
        S[j - 1] = (uint8_t)acc;
        any |= acc;
    }
    return any == 0;
}

// sym[0] = highest-degree symbol; corrects in place.  Returns corrected symbols or -1.
P25_FN int p25_rs_decode(const P25DevTables& T, uint8_t* sym, int n, int k) {
    const int nroots = n - k, t = nroots / 2;
    uint8_t S[16];
    if (p25_rs_syndromes(T, sym, n, nroots, S)) return 0;
    uint8_t lam[17];
    int L;
    if (nroots == 16)
        L = p25_berlekamp_massey<16>(T, S, lam);
    else if (nroots == 12)
        L = p25_berlekamp_massey<12>(T, S, lam);
    else
        L = p25_berlekamp_massey<8>(T, S, lam);
    if (L > t) return -1;
    uint8_t omega[16];
    for (int i = 0; i < nroots; i++) {
        int acc = 0;
        for (int j = 0; j <= i && j <= L; j++) acc ^= p25_gf_mul(T, lam[j], S[i - j]);
        omega[i] = (uint8_t)acc;
    }
    uint8_t dlam[17];
    for (int i = 0; i < 17; i++) dlam[i] = 0;
    for (int i = 1; i <= L; i += 2) dlam[i - 1] = lam[i];
    uint8_t loc[8], mag[8];
    int roots = 0;
    for (int p = 0; p < 63; p++) {
        const int xinv = T.gf_exp[(63 - p) % 63];
        if (p25_poly_eval(T, lam, L, xinv) != 0) continue;
        if (p >= n || roots >= t) return -1;
        const int den = p25_poly_eval(T, dlam, L > 0 ? L - 1 : 0, xinv);
        if (den == 0) return -1;
        const int mg = p25_gf_div(T, p25_poly_eval(T, omega, nroots - 1, xinv), den);
        if (mg == 0) return -1;
        loc[roots] = (uint8_t)(n - 1 - p);
        mag[roots] = (uint8_t)mg;
        roots++;
    }
    if (roots != L) return -1;
    for (int i = 0; i < roots; i++) sym[loc[i]] ^= mag[i];
    if (!p25_rs_syndromes(T, sym, n, nroots, S)) {
        for (int i = 0; i < roots; i++) sym[loc[i]] ^= mag[i];  // leave the word as received
        return -1;
    }
    return L;
}

// ------------------------------------------------------------------ half-rate trellis (TSBK)

// dibits: 98 received dibits (one per byte).  out12: decoded block.  Returns corrected bits or -1.
// Add-compare-select over the 4 states; survivor decisions are kept as 2 bits per state per step.
P25_FN int p25_trellis_half_decode(const P25DevTables& T, const uint8_t* dibits, uint8_t* out12) {
    int m0 = 0, m1 = 1 << 20, m2 = 1 << 20, m3 = 1 << 20;
    uint8_t from[49];  // 4 x 2-bit predecessor per step
    for (int i = 0; i < 49; i++) {
        const int slot = T.interleave[i];
        const int sym = (dibits[2 * slot] << 2) | dibits[2 * slot + 1];
        int nm[4];
        int packed = 0;
#pragma unroll
        for (int ns = 0; ns < 4; ns++) {
            int best = m0 + P25_POPC(T.trellis_pair[0 + ns] ^ sym), arg = 0;
            int c = m1 + P25_POPC(T.trellis_pair[4 + ns] ^ sym);
            if (c < best) { best = c; arg = 1; }
            c = m2 + P25_POPC(T.trellis_pair[8 + ns] ^ sym);
            if (c < best) { best = c; arg = 2; }
            c = m3 + P25_POPC(T.trellis_pair[12 + ns] ^ sym);
            if (c < best) { best = c; arg = 3; }
            nm[ns] = best;
            packed |= arg << (2 * ns);
        }
        m0 = nm[0]; m1 = nm[1]; m2 = nm[2]; m3 = nm[3];
        from[i] = (uint8_t)packed;
    }
    if (m0 > P25_VITERBI_MAX_FIX) return -1;
    for (int i = 0; i < 12; i++) out12[i] = 0;
    int st = 0;
    for (int i = 48; i >= 0; i--) {
        if (i < 48) out12[i >> 2] |= (uint8_t)(st << (6 - 2 * (i & 3)));
        st = (from[i] >> (2 * st)) & 3;
    }
    return m0;
}

// ------------------------------------------------------------------ 3/4-rate trellis (confirmed packet data blocks)
// dibits: 98 received dibits.  out18: decoded block (48 tribits, MSB first).  Returns corrected bits or -1.
// Eight path metrics in registers; the survivor of every state is kept as 3 bits of one word per step.
P25_FN int p25_trellis_34_decode(const P25DevTables& T, const uint8_t* dibits, uint8_t* out18) {
    int m[8];
    m[0] = 0;
#pragma unroll
    for (int s = 1; s < 8; s++) m[s] = 1 << 20;
    uint32_t from[49];
    for (int i = 0; i < 49; i++) {
        const int slot = T.interleave[i];
        const int sym = (dibits[2 * slot] << 2) | dibits[2 * slot + 1];
        int nm[8];
        uint32_t packed = 0;
#pragma unroll
        for (int ns = 0; ns < 8; ns++) {
            int key = ((m[0] + P25_POPC(T.trellis34_pair[ns] ^ sym)) << 3);
#pragma unroll
            for (int ps = 1; ps < 8; ps++) {
                const int k2 = ((m[ps] + P25_POPC(T.trellis34_pair[8 * ps + ns] ^ sym)) << 3) | ps;
                key = k2 < key ? k2 : key;      // equal metrics: the lower predecessor has the smaller key
            }
            nm[ns] = key >> 3;
            packed |= (uint32_t)(key & 7) << (3 * ns);
        }
#pragma unroll
        for (int s = 0; s < 8; s++) m[s] = nm[s];
        from[i] = packed;
    }
    if (m[0] > P25_VITERBI34_MAX_FIX) return -1;
    for (int i = 0; i < P25_PDU_BLOCK34_BYTES; i++) out18[i] = 0;
    int st = 0;
    for (int i = 48; i >= 0; i--) {
        if (i < 48) {
#pragma unroll
            for (int b = 0; b < 3; b++) {
                const int bit = 3 * i + b;
                out18[bit >> 3] |= (uint8_t)(((st >> (2 - b)) & 1) << (7 - (bit & 7)));
            }
        }
        st = (from[i] >> (3 * st)) & 7;
    }
    return m[0];
}

// CRC-CCITT of the P25 data blocks (TSBK, PDU header): x^16 + x^12 + x^5 + 1, zero start, inverted result
P25_FN uint32_t p25_crc_ccitt(const uint8_t* data, int n) {
    uint32_t crc = 0;
    for (int i = 0; i < n; i++) {
        crc ^= (uint32_t)data[i] << 8;
        P25_ROLLED
        for (int b = 0; b < 8; b++) crc = (crc & 0x8000) ? ((crc << 1) ^ 0x1021) & 0xFFFF : (crc << 1) & 0xFFFF;
    }
    return crc ^ 0xFFFF;
}

// ------------------------------------------------------------------ IMBE voice frame
// dibits: 72 received dibits.  chunks: u0..u7.  errors: 4 Golay + 3 Hamming corrected-bit counts.
P25_FN void p25_imbe_decode(const P25DevTables& T, const uint8_t* dibits, uint32_t* chunks, uint32_t* errors) {
    uint32_t cw[8];
    for (int c = 0; c < 8; c++) cw[c] = 0;
    for (int i = 0; i < 144; i++) {
        const uint32_t bit = (dibits[i >> 1] >> (1 - (i & 1))) & 1u;
        cw[T.imbe_cw[i]] |= bit << T.imbe_bit[i];
    }
    errors[0] = (uint32_t)p25_golay23_decode(T, cw[0], &chunks[0]);
    uint32_t p = (16u * chunks[0]) & 0xFFFF;
    for (int c = 1; c < 7; c++) {
        uint32_t mask = 0;
        const int nb = T.imbe_cw_bits[c];
        for (int b = 0; b < nb; b++) {
            p = (173u * p + 13849u) & 0xFFFF;
            mask = (mask << 1) | (p >> 15);
        }
        const uint32_t w = cw[c] ^ mask;
        if (c < 4)
            errors[c] = (uint32_t)p25_golay23_decode(T, w, &chunks[c]);
        else
            errors[c] = (uint32_t)p25_hamming15_decode(T, w, &chunks[c]);
    }
    chunks[7] = cw[7] & 0x7F;
}

// ------------------------------------------------------------------ bit packing helpers
P25_FN uint32_t p25_take_bits(const uint8_t* dibits, int bit0, int nbits) {
    uint32_t v = 0;
    P25_ROLLED
    for (int i = 0; i < nbits; i++) {
        const int b = bit0 + i;
        v = (v << 1) | ((dibits[b >> 1] >> (1 - (b & 1))) & 1u);
    }
    return v;
}

P25_FN void p25_pack_hexbits(const uint8_t* hex, int nhex, uint8_t* out) {
    const int nbytes = nhex * 6 / 8;
    for (int i = 0; i < nbytes; i++) out[i] = 0;
    for (int i = 0; i < nhex * 6; i++) {
        const int bit = (hex[i / 6] >> (5 - i % 6)) & 1;
        out[i >> 3] |= (uint8_t)(bit << (7 - (i & 7)));
    }
}
