// ORACLE -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.  See p25_oracle.hpp.
#include "p25_oracle.hpp"

#include <algorithm>
#include <atomic>
#include <thread>

namespace p25o {

// ===========================================================================
// GF(2^6) helpers  [STD primitive polynomial x^6+x+1]
// ===========================================================================
int gf_mul(int a, int b) {
    if (a == 0 || b == 0) return 0;
    return P25_GF_EXP[P25_GF_LOG[a] + P25_GF_LOG[b]];
}
int gf_div(int a, int b) {
    if (a == 0) return 0;
    return P25_GF_EXP[(P25_GF_LOG[a] + 63 - P25_GF_LOG[b]) % 63];
}
static inline int gf_pow_alpha(int e) { return P25_GF_EXP[((e % 63) + 63) % 63]; }

// Berlekamp-Massey over GF(64).  S[0..n-1] are the syndromes S_1..S_n.
// lambda must hold n+1 entries.  Returns the LFSR length L.
static int berlekamp_massey(const uint8_t* S, int n, uint8_t* lambda) {
    uint8_t C[40] = {0}, B[40] = {0}, T[40];
    C[0] = 1;
    B[0] = 1;
    int L = 0, m = 1, b = 1;
    for (int r = 0; r < n; r++) {
        int d = S[r];
        for (int i = 1; i <= L; i++) d ^= gf_mul(C[i], S[r - i]);
        if (d == 0) {
            m++;
        } else {
            int coef = gf_div(d, b);
            if (2 * L <= r) {
                std::memcpy(T, C, sizeof T);
                for (int i = 0; i + m <= n; i++) C[i + m] ^= gf_mul(coef, B[i]);
                L = r + 1 - L;
                std::memcpy(B, T, sizeof B);
                b = d;
                m = 1;
            } else {
                for (int i = 0; i + m <= n; i++) C[i + m] ^= gf_mul(coef, B[i]);
                m++;
            }
        }
    }
    for (int i = 0; i <= n; i++) lambda[i] = C[i];
    return L;
}

static inline int poly_eval(const uint8_t* p, int deg, int x) {
    int acc = 0;
    for (int i = deg; i >= 0; i--) acc = gf_mul(acc, x) ^ p[i];
    return acc;
}

// ===========================================================================
// BCH(63,16,23)  [STD]; surfaces as PacketNID, reference src/recv.rs:216-222
// ===========================================================================
static void bch_syndromes(uint64_t w, uint8_t* S) {
    for (int j = 1; j <= 2 * P25_BCH_T; j++) {
        int acc = 0;
        for (int i = 0; i < 63; i++)
            if ((w >> i) & 1) acc ^= gf_pow_alpha(j * i);
        S[j - 1] = (uint8_t)acc;
    }
}

bool bch_decode(uint64_t word63, uint16_t* data, int* nerr) {
    uint64_t w = word63 & 0x7FFFFFFFFFFFFFFFULL;
    uint8_t S[2 * P25_BCH_T];
    bch_syndromes(w, S);
    bool clean = true;
    for (int i = 0; i < 2 * P25_BCH_T; i++) clean &= (S[i] == 0);
    int fixed = 0;
    if (!clean) {
        uint8_t lam[2 * P25_BCH_T + 1];
        int L = berlekamp_massey(S, 2 * P25_BCH_T, lam);
        if (L > P25_BCH_T) return false;
        int roots = 0;
        uint64_t flip = 0;
        for (int p = 0; p < 63; p++) {
            if (poly_eval(lam, L, gf_pow_alpha(-p)) == 0) {
                flip |= 1ULL << p;
                roots++;
            }
        }
        if (roots != L) return false;
        w ^= flip;
        bch_syndromes(w, S);
        for (int i = 0; i < 2 * P25_BCH_T; i++)
            if (S[i]) return false;
        fixed = L;
    }
    *data = (uint16_t)(w >> 47);
    *nerr = fixed;
    return true;
}

// ===========================================================================
// Golay / Hamming / cyclic  [STD]
// ===========================================================================
static inline uint32_t polymod2(uint32_t a, uint32_t g, int gdeg, int abits) {
    for (int i = abits - 1; i >= gdeg; i--)
        if ((a >> i) & 1) a ^= g << (i - gdeg);
    return a;
}

int golay23_decode(uint32_t word, uint32_t* data12) {
    word &= 0x7FFFFF;
    uint32_t s = polymod2(word, P25_GOLAY_GEN, 11, 23);
    uint32_t e = P25_GOLAY23_SYN[s];
    word ^= e;
    *data12 = word >> 11;
    return __builtin_popcount(e);
}

int golay24_decode(uint32_t word, uint32_t* data12) {
    word &= 0xFFFFFF;
    uint32_t w23 = word >> 1;
    uint32_t s = polymod2(w23, P25_GOLAY_GEN, 11, 23);
    uint32_t e = P25_GOLAY23_SYN[s];
    uint32_t c23 = w23 ^ e;
    int nerr = __builtin_popcount(e);
    if ((uint32_t)(__builtin_popcount(c23) & 1) != (word & 1)) nerr++;
    *data12 = c23 >> 11;
    return nerr > 3 ? -1 : nerr;
}

int golay18_decode(uint32_t word, uint32_t* data6) {
    uint32_t d12;
    int n = golay24_decode(word & 0x3FFFF, &d12);
    *data6 = d12 & 0x3F;
    if (n < 0 || (d12 >> 6)) {
        *data6 = (word >> 12) & 0x3F;
        return -1;
    }
    return n;
}

int hamming15_decode(uint32_t word, uint32_t* data11) {
    word &= 0x7FFF;
    uint32_t p = 0;
    for (int i = 0; i < 11; i++)
        if ((word >> (14 - i)) & 1) p ^= P25_HAMMING15_COLS[i];
    uint32_t s = p ^ (word & 0xF);
    word ^= P25_HAMMING15_SYN[s];
    *data11 = word >> 4;
    return s ? 1 : 0;
}

int hamming10_decode(uint32_t word, uint32_t* data6) {
    word &= 0x3FF;
    uint32_t p = 0;
    for (int i = 0; i < 6; i++)
        if ((word >> (9 - i)) & 1) p ^= P25_HAMMING10_COLS[i];
    uint32_t s = p ^ (word & 0xF);
    uint32_t flip = P25_HAMMING10_SYN[s];
    if (flip == 0xFFFF) {
        *data6 = word >> 4;
        return -1;
    }
    word ^= flip;
    *data6 = word >> 4;
    return s ? 1 : 0;
}

int cyclic16_decode(uint32_t word, uint32_t* data8) {
    word &= 0xFFFF;
    uint32_t s = polymod2(word, P25_CYCLIC_GEN, 8, 16);
    uint32_t e = P25_CYCLIC16_SYN[s];
    if (e == 0xFFFF) {
        *data8 = word >> 8;
        return -1;
    }
    word ^= e;
    *data8 = word >> 8;
    return __builtin_popcount(e);
}

// ===========================================================================
// Reed-Solomon over GF(64)  [STD RS(24,12,13) RS(24,16,9) RS(36,20,17)]
// ===========================================================================
static bool rs_syndromes(const uint8_t* sym, int n, int nroots, uint8_t* S) {
    bool clean = true;
    for (int j = 1; j <= nroots; j++) {
        int a = gf_pow_alpha(j), acc = 0;
        for (int i = 0; i < n; i++) acc = gf_mul(acc, a) ^ sym[i];
        S[j - 1] = (uint8_t)acc;
        clean &= (acc == 0);
    }
    return clean;
}

int rs_decode(uint8_t* sym, int n, int k) {
    const int nroots = n - k, t = nroots / 2;
    uint8_t S[16];
    if (rs_syndromes(sym, n, nroots, S)) return 0;
    uint8_t lam[17];
    int L = berlekamp_massey(S, nroots, lam);
    if (L > t) return -1;
    // omega = S(x) * lambda(x) mod x^nroots
    uint8_t omega[16];
    for (int i = 0; i < nroots; i++) {
        int acc = 0;
        for (int j = 0; j <= i && j <= L; j++) acc ^= gf_mul(lam[j], S[i - j]);
        omega[i] = (uint8_t)acc;
    }
    // formal derivative of lambda (characteristic 2: odd-degree terms survive)
    uint8_t dlam[17] = {0};
    for (int i = 1; i <= L; i += 2) dlam[i - 1] = lam[i];
    uint8_t fixed[36];
    std::memcpy(fixed, sym, n);
    int roots = 0;
    for (int p = 0; p < 63; p++) {
        int xinv = gf_pow_alpha(-p);
        if (poly_eval(lam, L, xinv) != 0) continue;
        if (p >= n) return -1;  // error located in the shortened (always-zero) part
        int den = poly_eval(dlam, L > 0 ? L - 1 : 0, xinv);
        if (den == 0) return -1;
        int mag = gf_div(poly_eval(omega, nroots - 1, xinv), den);
        if (mag == 0) return -1;
        fixed[n - 1 - p] ^= (uint8_t)mag;
        roots++;
    }
    if (roots != L) return -1;
    if (!rs_syndromes(fixed, n, nroots, S)) return -1;
    std::memcpy(sym, fixed, n);
    return L;
}

// ===========================================================================
// Half-rate trellis (TSBK)  [STD]; surfaces as TrunkingControl, src/recv.rs:231
// ===========================================================================

int trellis_half_decode(const uint8_t* dibits98, uint8_t* out12) {
    uint8_t sym[49];
    for (int i = 0; i < 49; i++) {
        int slot = P25_INTERLEAVE[i];
        sym[i] = (uint8_t)((dibits98[2 * slot] << 2) | dibits98[2 * slot + 1]);
    }
    const int INF = 1 << 20;
    int m[4] = {0, INF, INF, INF};
    uint8_t from[49][4];
    for (int i = 0; i < 49; i++) {
        int nm[4];
        for (int ns = 0; ns < 4; ns++) {
            int best = INF * 2, arg = 0;
            for (int ps = 0; ps < 4; ps++) {
                int exp = P25_CONSTELLATION[P25_TRELLIS_HALF[ps * 4 + ns]];
                int cost = m[ps] + __builtin_popcount(exp ^ sym[i]);
                if (cost < best) {  // ties keep the lowest predecessor state
                    best = cost;
                    arg = ps;
                }
            }
            nm[ns] = best;
            from[i][ns] = (uint8_t)arg;
        }
        std::memcpy(m, nm, sizeof m);
    }
    // the flush dibit forces the encoder into state 0: only that survivor is legal
    if (m[0] > P25_VITERBI_MAX_FIX) return -1;
    uint8_t in[49];
    int st = 0;
    for (int i = 48; i >= 0; i--) {
        in[i] = (uint8_t)st;
        st = from[i][st];
    }
    std::memset(out12, 0, 12);
    for (int i = 0; i < 48; i++) out12[i / 4] |= (uint8_t)(in[i] << (6 - 2 * (i % 4)));
    return m[0];
}

// ===========================================================================
// 3/4-rate trellis (confirmed packet data blocks)  [STD]; surfaces only as the viterbiTribit stats family
// (src/hub.rs:570) and P25Error -- no MessageEvent variant carries packet data (src/recv.rs:214-233).
// 8 states (the previous tribit), 49 steps (48 tribits + the flush tribit), same interleaver as the 1/2-rate code.
// ===========================================================================
int trellis_34_decode(const uint8_t* dibits98, uint8_t* out18) {
    uint8_t sym[49];
    for (int i = 0; i < 49; i++) {
        const int slot = P25_INTERLEAVE[i];
        sym[i] = (uint8_t)((dibits98[2 * slot] << 2) | dibits98[2 * slot + 1]);
    }
    const int INF = 1 << 20;
    int m[8] = {0, INF, INF, INF, INF, INF, INF, INF};
    uint8_t from[49][8];
    for (int i = 0; i < 49; i++) {
        int nm[8];
        for (int ns = 0; ns < 8; ns++) {
            int best = 2 * INF, arg = 0;
            for (int ps = 0; ps < 8; ps++) {
                const int expect = P25_CONSTELLATION[P25_TRELLIS_3_4[ps * 8 + ns]];
                const int cost = m[ps] + __builtin_popcount(expect ^ sym[i]);
                if (cost < best) {  // ties keep the lowest predecessor state
                    best = cost;
                    arg = ps;
                }
            }
            nm[ns] = best;
            from[i][ns] = (uint8_t)arg;
        }
        std::memcpy(m, nm, sizeof m);
    }
    if (m[0] > P25_VITERBI34_MAX_FIX) return -1;  // the flush tribit forces state 0
    uint8_t in[49];
    int st = 0;
    for (int i = 48; i >= 0; i--) {
        in[i] = (uint8_t)st;
        st = from[i][st];
    }
    std::memset(out18, 0, 18);
    for (int i = 0; i < 48; i++)
        for (int b = 0; b < 3; b++) {
            const int bit = 3 * i + b;
            out18[bit >> 3] |= (uint8_t)(((in[i] >> (2 - b)) & 1) << (7 - (bit & 7)));
        }
    return m[0];
}

// ===========================================================================
// IMBE voice frame  [STD]; surfaces as VoiceFrame{chunks,errors}, src/audio.rs:76
// ===========================================================================
void imbe_decode(const uint8_t* dibits72, uint32_t chunks[8], uint32_t errors[7]) {
    uint32_t cw[8] = {0};
    for (int i = 0; i < 144; i++) {
        int bit = (dibits72[i / 2] >> (1 - (i & 1))) & 1;
        cw[P25_IMBE_SCHED_CW[i]] |= (uint32_t)bit << P25_IMBE_SCHED_BIT[i];
    }
    errors[0] = (uint32_t)golay23_decode(cw[0], &chunks[0]);
    uint32_t p = (16u * chunks[0]) & 0xFFFF;
    for (int c = 1; c < 7; c++) {
        uint32_t mask = 0;
        for (int b = 0; b < P25_IMBE_CW_BITS[c]; b++) {
            p = (173u * p + 13849u) & 0xFFFF;
            mask = (mask << 1) | (p >> 15);
        }
        cw[c] ^= mask;
        if (c < 4)
            errors[c] = (uint32_t)golay23_decode(cw[c], &chunks[c]);
        else
            errors[c] = (uint32_t)hamming15_decode(cw[c], &chunks[c]);
    }
    chunks[7] = cw[7] & 0x7F;
}

uint16_t crc_ccitt_p25(const uint8_t* data, int n) {
    uint32_t crc = 0;
    for (int i = 0; i < n; i++) {
        crc ^= (uint32_t)data[i] << 8;
        for (int b = 0; b < 8; b++) crc = (crc & 0x8000) ? ((crc << 1) ^ 0x1021) & 0xFFFF : (crc << 1) & 0xFFFF;
    }
    return (uint16_t)(crc ^ 0xFFFF);
}

// ===========================================================================
// Demod chain -- reference src/demod.rs:82-114
// ===========================================================================
size_t DemodChain::feed(const void* iq, size_t n, float* out, float* power_dbm) {
    size_t nout = 0;
    float pacc = 0.f;
    const uint8_t* u8 = (const uint8_t*)iq;
    const cf32* cf = (const cf32*)iq;
    for (size_t i = 0; i < n; i++) {
        cf32 s;
        if (fmt_ == FMT_U8)
            s = cf32{P25_IQ_LUT[u8[2 * i]], P25_IQ_LUT[u8[2 * i + 1]]};  // src/demod.rs:82-84
        else
            s = cf[i];
        if (use_front_ && !front_.feed(s, &s)) continue;
        cf32 d = s;
        if (!skip_decim_ && !decim_.feed(s, &d)) continue;        // src/demod.rs:87
        cf32 c = chan_.feed(d);                   // src/demod.rs:93
        pacc += c.re * c.re + c.im * c.im;        // src/demod.rs:125-127
        float f = fm_.feed(c);                    // src/demod.rs:109-111
        out[nout++] = avg_.feed(f);               // src/demod.rs:114
    }
    if (power_dbm) {
        float avg = nout ? pacc / (float)nout : 0.f;
        *power_dbm = 30.0f + 10.0f * std::log10(avg);  // src/demod.rs:129-133, R = 1
    }
    return nout;
}

// ===========================================================================
// MessageReceiver
// ===========================================================================
static bool g_always_correlate = true;

MessageReceiver::MessageReceiver() : n_(0) {
    std::memset(hist_, 0, sizeof hist_);
    std::memset(buf_, 0, sizeof buf_);
    std::memset(hex_, 0, sizeof hex_);
    pth_ = mid_ = nth_ = 0.f;
    next_sym_ = 0;
    frame_pos_ = 0;
    nid_bits_ = 0;
    nid_cnt_ = 0;
    duid_ = 0;
    cnt_ = blocks_ = part_ = chunks_ = 0;
    enter_sync();
}

void MessageReceiver::enter_sync() {
    state_ = ST_SYNC;
    det_have_prev_ = false;
    det_prev_above_ = false;
    det_prev_corr_ = 0.f;
}

void MessageReceiver::resync() { enter_sync(); }  // src/recv.rs:136,:179

float MessageReceiver::sample_at(int64_t idx) const {
    if (idx < 0) return 0.f;
    return hist_[idx & 255];
}

// Slicer thresholds from the sync word just matched [RECALL p25.rs sync.rs shape]:
// mean level of the +3 and -3 sync symbols, decision boundaries at mid and mid +- 2/3.
void MessageReceiver::lock(uint64_t idx) {
    const int64_t pk = (int64_t)idx - 1;  // correlation peaked on the previous sample
    float ps = 0.f, ns = 0.f;
    for (int i = 0; i < P25_FS_DIBITS; i++) {
        float v = sample_at(pk - (P25_FP_LEN - 1) + (int64_t)P25_SPS * i);
        if ((P25_SYNC_POS_MASK >> i) & 1)
            ps += v;
        else
            ns += v;
    }
    const float pavg = ps / 11.0f, navg = ns / 13.0f;
    mid_ = (pavg + navg) * 0.5f;
    pth_ = mid_ + (pavg - mid_) * (2.0f / 3.0f);
    nth_ = mid_ + (navg - mid_) * (2.0f / 3.0f);
    next_sym_ = (uint64_t)(pk + P25_SPS);
    frame_pos_ = P25_FS_DIBITS;
    state_ = ST_NID;
    nid_bits_ = 0;
    nid_cnt_ = 0;
}

int MessageReceiver::decide(float s) const {  // [STD] 01 +3, 00 +1, 10 -1, 11 -3
    if (s > pth_) return 1;
    if (s > mid_) return 0;
    if (s > nth_) return 2;
    return 3;
}

bool MessageReceiver::feed(float s, Event* ev) {
    const uint64_t idx = n_++;
    hist_[idx & 255] = s;
    if (state_ == ST_SYNC || g_always_correlate) {
        // frame-sync correlator, evaluated on every sample [RECALL p25.rs DataUnitReceiver]
        float corr = 0.f, energy = 0.f;
        const uint64_t base = idx - (P25_FP_LEN - 1);
        for (int k = 0; k < P25_FP_LEN; k++) {
            const float x = ((int64_t)(base + k) < 0) ? 0.f : hist_[(base + k) & 255];
            corr = std::fmaf(P25_SYNC_FP[k], x, corr);
            energy = std::fmaf(x, x, energy);
        }
        if (state_ == ST_SYNC) {
            const bool above = corr > 0.f && corr * corr >= P25_SYNC_RHO2_EFP * energy;
            const bool fire = det_have_prev_ && above && det_prev_above_ && corr <= det_prev_corr_;
            det_have_prev_ = true;
            det_prev_above_ = above;
            det_prev_corr_ = corr;
            if (fire) lock(idx);
            return false;
        }
    }
    if (idx != next_sym_) return false;
    next_sym_ += P25_SPS;
    return on_dibit(decide(s), idx, ev);
}

static void fill(Event* ev, uint32_t kind, uint64_t idx, const void* payload, uint32_t len) {
    ev->stream = 0;
    ev->kind = kind;
    ev->sample = idx;
    ev->len = len;
    std::memset(ev->payload, 0, sizeof ev->payload);
    std::memcpy(ev->payload, payload, len);
}

bool MessageReceiver::fail(uint32_t code, uint64_t idx, Event* ev) {
    fill(ev, EV_ERROR, idx, &code, 4);
    enter_sync();
    return true;
}

bool MessageReceiver::on_dibit(int d, uint64_t idx, Event* ev) {
    const uint32_t pos = frame_pos_++;
    if (pos % P25_STATUS_PERIOD == P25_STATUS_PERIOD - 1) {  // status symbol [STD]
        if (state_ == ST_FLUSH) enter_sync();
        return false;
    }
    switch (state_) {
        case ST_NID:
            nid_bits_ = (nid_bits_ << 2) | (uint64_t)d;
            if (++nid_cnt_ == P25_NID_DIBITS) return on_nid(idx, ev);
            return false;
        case ST_PAYLOAD:
            buf_[cnt_++] = (uint8_t)d;
            return on_payload(idx, ev);
        default:
            return false;
    }
}

bool MessageReceiver::on_nid(uint64_t idx, Event* ev) {
    uint16_t data;
    int nerr;
    if (!bch_decode(nid_bits_ >> 1, &data, &nerr)) {  // 64th bit (parity) is not used
        stats.bad(ST_BCH);
        return fail(ERR_BCH, idx, ev);
    }
    stats.ok(ST_BCH, (unsigned)nerr);
    duid_ = data & 0xF;
    cnt_ = blocks_ = part_ = chunks_ = 0;
    switch (duid_) {
        case 0x0: case 0x5: case 0xA: case 0xF: case 0x7:
            state_ = ST_PAYLOAD;
            break;
        case 0x3:
            state_ = ST_FLUSH;
            break;
        case 0xC:  // packet data: header block, then the data blocks it announces; none of them yields a
            state_ = ST_PAYLOAD;  // MessageEvent (src/recv.rs:214-233), only stats and errors
            break;
        default:
            return fail(ERR_UNKNOWN_NID, idx, ev);
    }
    uint8_t p[3] = {(uint8_t)((data >> 4) & 0xFF), (uint8_t)(data >> 12), (uint8_t)duid_};
    fill(ev, EV_NID, idx, p, 3);
    return true;
}

// pack `nbits` bits taken MSB-first from a dibit array starting at bit offset `bit0`
static uint32_t take_bits(const uint8_t* dibits, int bit0, int nbits) {
    uint32_t v = 0;
    for (int i = 0; i < nbits; i++) {
        int b = bit0 + i;
        v = (v << 1) | ((dibits[b >> 1] >> (1 - (b & 1))) & 1u);
    }
    return v;
}

static void pack_hexbits(const uint8_t* hex, int nhex, uint8_t* out) {
    int nbytes = nhex * 6 / 8;
    std::memset(out, 0, nbytes);
    for (int i = 0; i < nhex * 6; i++) {
        int bit = (hex[i / 6] >> (5 - i % 6)) & 1;
        out[i / 8] |= (uint8_t)(bit << (7 - i % 8));
    }
}

bool MessageReceiver::on_payload(uint64_t idx, Event* ev) {
    switch (duid_) {
        case 0x7: {  // TSDU: up to three TSBKs
            if (cnt_ < P25_TSBK_DIBITS) return false;
            uint8_t out[12];
            int fixed = trellis_half_decode(buf_, out);
            cnt_ = 0;
            if (fixed < 0) {
                stats.bad(ST_VITERBI_DIBIT);
                return fail(ERR_VITERBI_DIBIT, idx, ev);
            }
            stats.ok(ST_VITERBI_DIBIT, (unsigned)fixed);
            blocks_++;
            if ((out[0] & 0x80) || blocks_ == 3) state_ = ST_FLUSH;
            fill(ev, EV_TSBK, idx, out, 12);  // emitted before any CRC check (src/recv.rs:242)
            return true;
        }
        case 0xC: {  // PDU: 98-dibit blocks.  blocks_ = blocks decoded, part_ = data blocks announced, chunks_ = confirmed
            if (cnt_ < P25_TSBK_DIBITS) return false;
            cnt_ = 0;
            if (blocks_ == 0) {  // header block, always 1/2-rate [STD]
                uint8_t h[12];
                const int fixed = trellis_half_decode(buf_, h);
                if (fixed < 0) {
                    stats.bad(ST_VITERBI_DIBIT);
                    return fail(ERR_VITERBI_DIBIT, idx, ev);
                }
                stats.ok(ST_VITERBI_DIBIT, (unsigned)fixed);
                if (crc_ccitt_p25(h, 10) != (uint16_t)((h[10] << 8) | h[11])) {
                    enter_sync();  // the length field cannot be trusted: drop lock (no P25Error names a CRC)
                    return false;
                }
                blocks_ = 1;
                part_ = h[6] & 0x7F;
                chunks_ = (h[0] & 0x1F) == P25_PDU_FORMAT_CONFIRMED;
                if (part_ == 0) state_ = ST_FLUSH;
                return false;
            }
            if (chunks_) {  // confirmed data: 3/4-rate blocks
                uint8_t d[18];
                const int fixed = trellis_34_decode(buf_, d);
                if (fixed < 0) {
                    stats.bad(ST_VITERBI_TRIBIT);
                    return fail(ERR_VITERBI_TRIBIT, idx, ev);
                }
                stats.ok(ST_VITERBI_TRIBIT, (unsigned)fixed);
            } else {
                uint8_t d[12];
                const int fixed = trellis_half_decode(buf_, d);
                if (fixed < 0) {
                    stats.bad(ST_VITERBI_DIBIT);
                    return fail(ERR_VITERBI_DIBIT, idx, ev);
                }
                stats.ok(ST_VITERBI_DIBIT, (unsigned)fixed);
            }
            if (++blocks_ == 1 + part_) state_ = ST_FLUSH;
            return false;
        }
        case 0x0: {  // HDU
            if (cnt_ < P25_HDU_DIBITS) return false;
            for (int w = 0; w < 36; w++) {
                uint32_t d6;
                int n = golay18_decode(take_bits(buf_, 18 * w, 18), &d6);
                if (n < 0) stats.bad(ST_GOLAY_SHORT); else stats.ok(ST_GOLAY_SHORT, (unsigned)n);
                hex_[w] = (uint8_t)d6;
            }
            int n = rs_decode(hex_, 36, 20);
            if (n < 0) {
                stats.bad(ST_RS_LONG);
                return fail(ERR_RS, idx, ev);
            }
            stats.ok(ST_RS_LONG, (unsigned)n);
            uint8_t out[15];
            pack_hexbits(hex_, 20, out);
            state_ = ST_FLUSH;
            fill(ev, EV_VOICE_HEADER, idx, out, 15);
            return true;
        }
        case 0xF: {  // TDULC
            if (cnt_ < P25_TDULC_DIBITS) return false;
            for (int w = 0; w < 12; w++) {
                uint32_t d12;
                uint32_t word = take_bits(buf_, 24 * w, 24);
                int n = golay24_decode(word, &d12);
                if (n < 0) {
                    stats.bad(ST_GOLAY_EXT);
                    d12 = word >> 12;
                } else {
                    stats.ok(ST_GOLAY_EXT, (unsigned)n);
                }
                hex_[2 * w] = (uint8_t)(d12 >> 6);
                hex_[2 * w + 1] = (uint8_t)(d12 & 0x3F);
            }
            int n = rs_decode(hex_, 24, 12);
            if (n < 0) {
                stats.bad(ST_RS_SHORT);
                return fail(ERR_RS, idx, ev);
            }
            stats.ok(ST_RS_SHORT, (unsigned)n);
            uint8_t out[9];
            pack_hexbits(hex_, 12, out);
            state_ = ST_FLUSH;
            fill(ev, EV_VOICE_TERM, idx, out, 9);
            return true;
        }
        case 0x5:
        case 0xA: {  // LDU1 / LDU2
            const int start = P25_LDU_PART_START[part_], len = P25_LDU_PART_LEN[part_];
            if (cnt_ < start + len) return false;
            const int kind = P25_LDU_PART_KIND[part_];
            const uint8_t* d = buf_ + start;
            part_++;
            if (kind == 0) {  // IMBE voice frame
                uint32_t pl[15];
                imbe_decode(d, pl, pl + 8);
                for (int i = 0; i < 4; i++) stats.ok(ST_GOLAY_STD, pl[8 + i]);
                for (int i = 4; i < 7; i++) stats.ok(ST_HAMMING_STD, pl[8 + i]);
                if (part_ == P25_LDU_PARTS) state_ = ST_FLUSH;
                fill(ev, EV_VOICE_FRAME, idx, pl, 60);
                return true;
            }
            if (kind == 1) {  // four Hamming(10,6,3) words of link control / crypto sync
                for (int w = 0; w < 4; w++) {
                    uint32_t d6;
                    int n = hamming10_decode(take_bits(d, 10 * w, 10), &d6);
                    if (n < 0) stats.bad(ST_HAMMING_SHORT); else stats.ok(ST_HAMMING_SHORT, (unsigned)n);
                    hex_[4 * chunks_ + w] = (uint8_t)d6;
                }
                if (++chunks_ < 6) return false;
                const bool lc = duid_ == 0x5;
                int n = rs_decode(hex_, 24, lc ? 12 : 16);
                if (n < 0) {
                    stats.bad(lc ? ST_RS_SHORT : ST_RS_MED);
                    return fail(ERR_RS, idx, ev);
                }
                stats.ok(lc ? ST_RS_SHORT : ST_RS_MED, (unsigned)n);
                uint8_t out[12];
                pack_hexbits(hex_, lc ? 12 : 16, out);
                fill(ev, lc ? EV_LINK_CONTROL : EV_CRYPTO_CONTROL, idx, out, lc ? 9 : 12);
                return true;
            }
            // low speed data: two cyclic(16,8,5) words
            uint32_t frag = 0;
            for (int w = 0; w < 2; w++) {
                uint32_t d8;
                int n = cyclic16_decode(take_bits(d, 16 * w, 16), &d8);
                if (n < 0) stats.bad(ST_CYCLIC); else stats.ok(ST_CYCLIC, (unsigned)n);
                frag = (frag << 8) | d8;
            }
            fill(ev, EV_LSD, idx, &frag, 4);
            return true;
        }
        default:
            return false;
    }
}

}  // namespace p25o

// ===========================================================================
// C entry points (ctypes) -- used by tests/, smoke() and bench.py cpu legs only
// ===========================================================================
using namespace p25o;

extern "C" {

void p25o_set_always_correlate(int on) { g_always_correlate = on != 0; }

void* p25o_demod_new(int fmt, int mode) { return new DemodChain(fmt, mode); }
void p25o_demod_free(void* h) { delete (DemodChain*)h; }
size_t p25o_demod_feed(void* h, const void* iq, size_t n, float* out, float* power_dbm) {
    return ((DemodChain*)h)->feed(iq, n, out, power_dbm);
}

void* p25o_recv_new(void) { return new MessageReceiver(); }
void p25o_recv_free(void* h) { delete (MessageReceiver*)h; }
void p25o_recv_resync(void* h) { ((MessageReceiver*)h)->resync(); }
int p25o_recv_state(void* h) { return ((MessageReceiver*)h)->state(); }
// ReplayReceiver::feed loop (reference src/replay.rs:40-57): returns events appended.
size_t p25o_recv_feed(void* h, const float* s, size_t n, Event* out, size_t cap, uint32_t stream) {
    MessageReceiver* r = (MessageReceiver*)h;
    size_t ne = 0;
    Event ev;
    for (size_t i = 0; i < n; i++) {
        if (!r->feed(s[i], &ev)) continue;
        ev.stream = stream;
        if (ne < cap) out[ne] = ev;
        ne++;
    }
    return ne;
}
void p25o_recv_stats(void* h, uint64_t* out48, int clear) {
    MessageReceiver* r = (MessageReceiver*)h;
    for (int i = 0; i < ST_FAMILIES; i++) {
        out48[4 * i + 0] = r->stats.c[i].words;
        out48[4 * i + 1] = r->stats.c[i].errs;
        out48[4 * i + 2] = r->stats.c[i].size;
        out48[4 * i + 3] = r->stats.c[i].fixed;
    }
    if (clear) r->stats.clear();
}

// FEC unit entry points
int p25o_bch_decode(uint64_t w, uint16_t* data, int* nerr) { return bch_decode(w, data, nerr) ? 1 : 0; }
int p25o_golay23_decode(uint32_t w, uint32_t* d) { return golay23_decode(w, d); }
int p25o_golay24_decode(uint32_t w, uint32_t* d) { return golay24_decode(w, d); }
int p25o_golay18_decode(uint32_t w, uint32_t* d) { return golay18_decode(w, d); }
int p25o_hamming15_decode(uint32_t w, uint32_t* d) { return hamming15_decode(w, d); }
int p25o_hamming10_decode(uint32_t w, uint32_t* d) { return hamming10_decode(w, d); }
int p25o_cyclic16_decode(uint32_t w, uint32_t* d) { return cyclic16_decode(w, d); }
int p25o_rs_decode(uint8_t* sym, int n, int k) { return rs_decode(sym, n, k); }
int p25o_trellis_half_decode(const uint8_t* d98, uint8_t* out12) { return trellis_half_decode(d98, out12); }
int p25o_trellis_34_decode(const uint8_t* d98, uint8_t* out18) { return trellis_34_decode(d98, out18); }
void p25o_imbe_decode(const uint8_t* d72, uint32_t* chunks8, uint32_t* errors7) { imbe_decode(d72, chunks8, errors7); }
uint32_t p25o_crc_ccitt(const uint8_t* d, int n) { return crc_ccitt_p25(d, n); }

// Batched driver for the CPU baseline: S independent streams, each with its own
// DemodChain + MessageReceiver, spread over `threads` host threads.  Mirrors one
// reference process per stream (demod thread + receiver thread, src/main.rs:270-287).
// iq: [S][n] samples of the given format.  Returns total events; fills per-stream counts.
size_t p25o_batch_run(int fmt, int front, const void* iq, size_t n_streams, size_t n_per_stream, int threads,
                      Event* out, size_t cap_per_stream, uint32_t* counts) {
    std::atomic<size_t> next(0), total(0);
    const size_t bps = (fmt == FMT_U8 ? 2 : 8) * n_per_stream;
    auto work = [&]() {
        std::vector<float> bb(n_per_stream / (front ? 50 : 5) + 16);
        for (;;) {
            size_t s = next.fetch_add(1);
            if (s >= n_streams) break;
            DemodChain dc(fmt, front != 0);
            MessageReceiver rx;
            size_t nb = dc.feed((const uint8_t*)iq + s * bps, n_per_stream, bb.data(), nullptr);
            size_t ne = 0;
            Event ev;
            for (size_t i = 0; i < nb; i++) {
                if (!rx.feed(bb[i], &ev)) continue;
                ev.stream = (uint32_t)s;
                if (out && ne < cap_per_stream) out[s * cap_per_stream + ne] = ev;
                ne++;
            }
            if (counts) counts[s] = (uint32_t)ne;
            total += ne;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    return total.load();
}

}  // extern "C"
