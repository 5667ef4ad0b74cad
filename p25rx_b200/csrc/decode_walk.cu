// decode_walk.cu -- frame sync, symbol slicing, deframing and FEC decode for many streams.
//
// Replaces, per stream, the loop `for &s in samples { msg.feed(s) }` of the reference
// (src/recv.rs:148-150, :204-234; src/replay.rs:40-57) where `msg` is p25::MessageReceiver
// (crate p25 @a96c564, not vendored).  One warp owns one stream.  Instead of stepping one
// sample at a time the warp advances in bulk between "decision points":
//
//   SYNC    128 candidate positions per step, four consecutive ones per lane: the 231-tap frame-sync
//           correlation and the window energy of every position run on packed FFMA2 from a register
//           window fed by 16-byte shared-memory loads; the first position whose predicate fires is
//           found with a warp-wide minimum, the predecessor of a lane's first position by shuffle.
//   symbols 64 symbol instants per step (stride 10 samples); lanes slice their samples against the
//           three thresholds, a ballot removes status symbols (every 36th dibit) and gives each
//           data dibit its index in the unit buffer.
//   unit    when the dibit that completes a NID / TSBK block / voice frame / ... has been stored the
//           warp decodes it cooperatively (BCH syndromes and Chien search per lane, Viterbi on 16
//           lanes, IMBE / Golay / Hamming / cyclic words one per lane, Reed-Solomon with per-lane
//           syndromes and Chien/Forney); lane 0 alone touches the receiver state, the stats and
//           queues the event, which the warp writes out 20 lanes wide.
//
// The grid covers the streams in one pass, or -- beside the next chunk's demod kernel -- is a
// persistent grid of two CTAs per SM that walks them in several passes (p25cu_launch_walk).
//
// Every step is an exact restatement of the sample-at-a-time rule (oracle/p25_oracle.cpp),
// so events and their sample indices are bit-identical for any chunking.  This file is
// compiled with --fmad=false: the only fused operations are the explicit fmaf() of the
// correlator, in the same order as the oracle.
#include <mutex>

#include "p25cu_internal.cuh"

#define FULL 0xFFFFFFFFu

__constant__ float c_sync_fp[P25_FP_LEN];

#define SEARCH_N 128                          // candidate positions per search step (4 per lane)
#define WIN_LEN (P25_FP_LEN - 1 + SEARCH_N)  // samples needed to correlate them
#define WIN_PAD 364                           // staged window, padded to whole 16-byte groups past the last LDS.128
#define FPB_LEN 256                           // packed fingerprint pairs, operand of the tensor-pipe prefilter
// A search step may be skipped when no position can reach the detector's threshold corr > 0 && corr^2 >= rho^2 |fp|^2 en.
// The prefilter bounds both sides from the safe direction:
//   * the 128 correlations in BF16 on the tensor pipe: inputs rounded to 8 mantissa bits perturb a product by less than
//     2^-8 relative, the normalised correlation therefore by less than 2^-8 = 0.004 (Cauchy-Schwarz on the rounding terms);
//   * the window energy from BELOW: the 27 whole 8-sample blocks that lie inside every window of a row of 8 positions
//     (216 of the 231 samples), from block energies and one warp prefix sum.
// Testing corr' > 0 && corr'^2 >= (rho - 0.02)^2 |fp|^2 en_low can therefore only err towards running the exact correlator.
#ifndef P25_PREFILTER_QUIET
#define P25_PREFILTER_QUIET 0                 // empty search steps in a row before the prefilter is consulted (A/B: 0 is best where it is used)
#endif
#define P25_PREFILTER_RHO2_EFP (P25_SYNC_RHO2_EFP * (0.63f * 0.63f) / (0.65f * 0.65f))

struct PendingEvent {
    unsigned kind, len, valid, pad;
    unsigned long long idx;
};

struct WalkShared {
    P25DevTables T;
    WalkState ws[P25CU_WALK_WARPS];
    alignas(16) float win[P25CU_WALK_WARPS][2 * WIN_PAD];    // the search window twice: [0, WIN_PAD) and shifted by one sample
    PendingEvent pend[P25CU_WALK_WARPS];
    alignas(16) unsigned char scr[P25CU_WALK_WARPS][192];   // decoder work area (syndromes, locator, IMBE results)
    unsigned surv[P25CU_WALK_WARPS][52];        // 3/4-rate trellis survivors: 8 states x 3 bits per step
    unsigned short pn_a[116], pn_c[116];        // IMBE PN generator n steps ahead: p_n = pn_a[n] * p_0 + pn_c[n] (mod 2^16)
    unsigned char imbe_src[8 * 24];             // inverse of the IMBE interleave schedule: [code word][bit] -> frame bit
};

// packed pair of IEEE fused multiply-adds (one FFMA2): each half is exactly fmaf()
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(r)
        : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)),
          "l"(*reinterpret_cast<const unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&r);
}
static_assert(sizeof(P25DevTables) % 16 == 0, "tables are staged with 16-byte copies");
// pull the cache lines the next step will read into L1 (the row stays in L2 after the demod kernel wrote it)
__device__ __forceinline__ void prefetch_l1(const float* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

__device__ __forceinline__ void mma_bf16(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// two floats -> one register of two BF16 (round to nearest), lo in bits 0..15
__device__ __forceinline__ unsigned pack_bf16(float lo, float hi) {
    unsigned r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// Conservative test of one 128-position search step (see P25_PREFILTER_RHO2_EFP): can ANY position be above the
// detector's threshold?  The 128 correlations are one 16 x 8 tile of the Hankel product
//     corr[8 a + b] = sum_j win[8 a + j] * fp[j - b],   j = 0 .. 239 (fp = 0 outside 0 .. 230),
// fifteen m16n8k16 steps whose A fragments are 8-byte shared-memory loads of the staged window (rows overlap: no copy of
// the operand is built; the 32 lanes of a load cover 256 consecutive bytes) and whose B fragments come from a table of
// packed fingerprint pairs.  15 tensor instructions instead of 924 FFMA2 per lane; the exact correlator then only runs
// on steps that hold a real candidate.  Warp-uniform result.
__device__ __noinline__ bool sync_prefilter(const float* __restrict__ win, const unsigned* __restrict__ fpb, float* __restrict__ pe, int lane) {
    const int g = lane >> 2, t = lane & 3;
    // ---- lower bound of the window energies: blocks of 8 samples, prefix sums, 27 whole blocks per row of positions
    {
        const float4* w4 = reinterpret_cast<const float4*>(win);
        const float4 u0 = w4[2 * lane], u1 = w4[2 * lane + 1];
        float ea = u0.x * u0.x + u0.y * u0.y + u0.z * u0.z + u0.w * u0.w + u1.x * u1.x + u1.y * u1.y + u1.z * u1.z + u1.w * u1.w, eb = 0.f;
        if (lane < 13) {                          // blocks 32 .. 44: samples 256 .. 359
            const float4 v0 = w4[2 * (lane + 32)], v1 = w4[2 * (lane + 32) + 1];
            eb = v0.x * v0.x + v0.y * v0.y + v0.z * v0.z + v0.w * v0.w + v1.x * v1.x + v1.y * v1.y + v1.z * v1.z + v1.w * v1.w;
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float xa = __shfl_up_sync(FULL, ea, o), xb = __shfl_up_sync(FULL, eb, o);
            if (lane >= o) {
                ea += xa;
                eb += xb;
            }
        }
        eb += __shfl_sync(FULL, ea, 31);
        pe[lane] = ea;                            // pe[m] = energy of blocks 0 .. m
        if (lane < 13) pe[32 + lane] = eb;
    }
    const float2* wa = reinterpret_cast<const float2*>(win + 8 * g + 2 * t);   // A[g][2t, 2t+1] of step 0; row g + 8 is 64 samples on
    const unsigned* fb = fpb + 8 + 2 * t - g;                                     // B[2t, 2t+1][g] of step 0: (fp[2t - g], fp[2t + 1 - g])
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    // rows g + 8 of step s are rows g of step s + 4 (the window moves 8 samples per row, 16 per step ... 64 = 4 steps):
    // every fragment register is loaded and packed once and used twice, 38 instead of 60 shared-memory loads per tile
    unsigned q0[19], q2[19];
#pragma unroll
    for (int s = 0; s < 19; s++) {
        if (s < 4) {
            const float2 x0 = wa[8 * s], x2 = wa[8 * s + 4];
            q0[s] = pack_bf16(x0.x, x0.y);
            q2[s] = pack_bf16(x2.x, x2.y);
        }
        if (s + 4 < 19) {
            const float2 x0 = wa[8 * (s + 4)], x2 = wa[8 * (s + 4) + 4];
            q0[s + 4] = pack_bf16(x0.x, x0.y);
            q2[s + 4] = pack_bf16(x2.x, x2.y);
        }
        if (s < 15) {
            const unsigned a[4] = {q0[s], q0[s + 4], q2[s], q2[s + 4]};
            mma_bf16(c, a, fb[16 * s], fb[16 * s + 8]);
        }
    }
    __syncwarp();
    const float l0 = pe[g + 27] - pe[g], l1 = pe[g + 35] - pe[g + 8];   // blocks a + 1 .. a + 27 lie inside every window of row a
    const bool hit = (c[0] > 0.f && c[0] * c[0] >= P25_PREFILTER_RHO2_EFP * l0) || (c[1] > 0.f && c[1] * c[1] >= P25_PREFILTER_RHO2_EFP * l0) ||
                     (c[2] > 0.f && c[2] * c[2] >= P25_PREFILTER_RHO2_EFP * l1) || (c[3] > 0.f && c[3] * c[3] >= P25_PREFILTER_RHO2_EFP * l1);
    return __any_sync(FULL, hit);
}

struct WarpCtx {
    const WalkParams* p;
    const P25DevTables* T;
    WalkState* ws;
    unsigned stream;
    PendingEvent* pend;       // event decided by lane 0 at this decision point (at most one), written out by the warp
    unsigned char* payload;   // its payload bytes (shared memory, 60 bytes)
};

// ---------------------------------------------------------------- lane-0 helpers
__device__ __forceinline__ void stat_ok(const WarpCtx& c, int fam, unsigned fixed) {
    unsigned long long* s = c.p->stats + ((size_t)c.stream * P25CU_ST_FAMILIES + fam) * 3;
    s[0] += 1;
    s[2] += fixed;
}
__device__ __forceinline__ void stat_bad(const WarpCtx& c, int fam) {
    unsigned long long* s = c.p->stats + ((size_t)c.stream * P25CU_ST_FAMILIES + fam) * 3;
    s[0] += 1;
    s[1] += 1;
}

// lane 0: queue the (single) event of this decision point; flush_event() writes it out with the whole warp
__device__ __forceinline__ void emit(const WarpCtx& c, unsigned kind, unsigned long long idx, const void* payload, unsigned len) {
    WalkState& ws = *c.ws;
    if (ws.n_events >= c.p->ev_cap) {
        ws.overflow = 1;
        return;
    }
    const unsigned char* src = (const unsigned char*)payload;
    if (src != c.payload)
        for (unsigned i = 0; i < len; i++) c.payload[i] = src[i];
    c.pend->kind = kind;
    c.pend->len = len;
    c.pend->idx = idx;
    c.pend->valid = 1;
}

// whole warp: one packed record = 3 header words + ceil(len / 4) payload words, one word per lane
// (word 0 stream, word 1 sample bits 0..31, word 2 sample bits 32..47 | kind << 16 | len << 24)
__device__ __forceinline__ void flush_event(const WarpCtx& c, int lane) {
    if (!c.pend->valid) return;
    WalkState& ws = *c.ws;
    const unsigned len = c.pend->len, nw = p25cu_packed_words(len);
    unsigned w = 0;
    if (lane == 0) w = c.stream;
    else if (lane == 1) w = (unsigned)c.pend->idx;
    else if (lane == 2) w = ((unsigned)(c.pend->idx >> 32) & 0xFFFFu) | (c.pend->kind << 16) | (len << 24);
    else if (lane < (int)nw) {
        const unsigned b0 = 4 * (lane - 3);
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (b0 + i < len) w |= (unsigned)c.payload[b0 + i] << (8 * i);
    }
    unsigned* dst = c.p->slots + (size_t)c.stream * c.p->ev_cap * P25CU_SLOT_WORDS + ws.n_words;
    if (lane < (int)nw) dst[lane] = w;
    __syncwarp();
    if (lane == 0) {
        ws.n_events++;
        ws.n_words += nw;
        c.pend->valid = 0;
    }
    __syncwarp();
}

__device__ __forceinline__ void enter_sync(WalkState& ws, unsigned long long pos) {
    ws.state = WS_SYNC;
    ws.pos = pos;
    ws.have_prev = 0;
    ws.prev_above = 0;
    ws.prev_corr = 0.f;
}

__device__ __forceinline__ void fail(const WarpCtx& c, unsigned code, unsigned long long idx) {
    emit(c, P25CU_EV_ERROR, idx, &code, 4);
    enter_sync(*c.ws, idx + 1);
}

// NID complete (32 dibits in ws.buf).  PacketNID event, reference src/recv.rs:216-222.  Called by the whole warp.
__device__ __forceinline__ void nid_apply(const WarpCtx& c, unsigned long long idx, int nerr, unsigned data) {  // lane 0
    WalkState& ws = *c.ws;
    if (nerr < 0) {
        stat_bad(c, P25CU_ST_BCH);
        fail(c, P25CU_E_BCH, idx);
        return;
    }
    stat_ok(c, P25CU_ST_BCH, (unsigned)nerr);
    ws.duid = (int)(data & 0xF);
    ws.cnt = ws.blocks = ws.part = ws.chunks = 0;
    switch (ws.duid) {
        case 0x0: case 0x5: case 0xA: case 0xF: case 0x7:
        case 0xC:   // packet data: header block + the data blocks it announces (stats and errors only, no MessageEvent)
            ws.state = WS_PAYLOAD;
            break;
        case 0x3:
            ws.state = WS_FLUSH;
            break;
        default:
            fail(c, P25CU_E_UNKNOWN_NID, idx);
            return;
    }
    const unsigned char pl[3] = {(unsigned char)((data >> 4) & 0xFF), (unsigned char)(data >> 12), (unsigned char)ws.duid};
    emit(c, P25CU_EV_NID, idx, pl, 3);
}

// ---------------------------------------------------------------- warp-cooperative decoders
// BCH(63,16,23), all 32 lanes: lanes 0..21 each sum one syndrome, lane 0 runs Berlekamp-Massey only when a
// syndrome is non-zero, the Chien search covers two positions per lane.  Same bounded-distance result as
// p25_bch_decode (the per-thread form used by the unit kernels).  Returns corrected bits or -1, uniformly.
__device__ __forceinline__ unsigned warp_bch_syndromes(const P25DevTables& T, unsigned long long w, int lane) {
    // S_j = r(alpha^j) over GF(64) is linear in the received bits: each of its six bits is the parity of the word under a
    // fixed mask (p25_fill_tables) -- 6 AND + POPC pairs per lane instead of a 63-step loop over the bits
    unsigned acc = 0;
    if (lane < 2 * P25_BCH_T) {
#pragma unroll
        for (int b = 0; b < 6; b++) acc |= (unsigned)(__popcll(w & T.bch_par[lane][b]) & 1) << b;
    }
    return acc;
}

__device__ __noinline__ int warp_bch_decode(const P25DevTables& T, unsigned char* scratch, unsigned long long word63, int lane,
                                            unsigned* data16) {
    unsigned long long w = word63 & 0x7FFFFFFFFFFFFFFFULL;
    const unsigned syn = warp_bch_syndromes(T, w, lane);
    int fixed = 0;
    if (__ballot_sync(FULL, syn != 0)) {
        if (lane < 2 * P25_BCH_T) scratch[lane] = (unsigned char)syn;
        __syncwarp();
        int L = 0;
        if (lane == 0) {
            unsigned char S[2 * P25_BCH_T], lam[2 * P25_BCH_T + 1];
            for (int i = 0; i < 2 * P25_BCH_T; i++) S[i] = scratch[i];
            L = p25_berlekamp_massey<2 * P25_BCH_T>(T, S, lam);
            for (int i = 0; i <= P25_BCH_T; i++) scratch[32 + i] = lam[i];
        }
        L = __shfl_sync(FULL, L, 0);
        __syncwarp();
        if (L > P25_BCH_T) return -1;
        unsigned char lam[P25_BCH_T + 1];
        for (int i = 0; i <= P25_BCH_T; i++) lam[i] = scratch[32 + i];
        const int p1 = lane, p2 = lane + 32;
        const bool r1 = p25_poly_eval(T, lam, L, T.gf_exp[(63 - p1) % 63]) == 0;
        const bool r2 = p2 < 63 && p25_poly_eval(T, lam, L, T.gf_exp[(63 - p2) % 63]) == 0;
        const unsigned lo = __ballot_sync(FULL, r1), hi = __ballot_sync(FULL, r2);
        if (__popc(lo) + __popc(hi) != L) return -1;
        w ^= (unsigned long long)lo | ((unsigned long long)hi << 32);
        if (__ballot_sync(FULL, warp_bch_syndromes(T, w, lane) != 0)) return -1;
        fixed = L;
    }
    *data16 = (unsigned)(w >> 47);
    return fixed;
}

// Half-rate trellis, 16 lanes = (next state, previous state) pairs; add-compare-select by two shuffle
// steps on the key (path metric << 2 | previous state), which keeps the lowest predecessor on ties exactly
// like p25_trellis_half_decode.  Survivors: one byte per step in scratch, traced back by lane 0.
__device__ __noinline__ int warp_trellis_half_decode(const P25DevTables& T, unsigned char* scratch, const unsigned char* dibits,
                                                     int lane, unsigned char* out12) {
    // deinterleave: lane holds received symbols lane and lane + 32
    int s0 = 0, s1 = 0;
    {
        const int slot = T.interleave[lane];
        s0 = (dibits[2 * slot] << 2) | dibits[2 * slot + 1];
        if (lane + 32 < 49) {
            const int slot1 = T.interleave[lane + 32];
            s1 = (dibits[2 * slot1] << 2) | dibits[2 * slot1 + 1];
        }
    }
    // Exact shortcut: the 16 (previous state, next state) transitions emit 16 distinct symbols, so every
    // received symbol names one transition.  If consecutive transitions chain up from state 0 to state 0 the
    // received block is a code word (metric 0) and, the free distance being 5, Viterbi would return exactly
    // this path.  Fully parallel; the add-compare-select loop below only runs for blocks with bit errors.
    {
        int t0i = 0, t1i = 0;   // transition index prev*4+next of symbols lane and lane+32
#pragma unroll
        for (int t = 0; t < 16; t++) {
            const int pr = T.trellis_pair[t];
            if (pr == s0) t0i = t;
            if (pr == s1) t1i = t;
        }
        const int nx0 = t0i & 3, nx1 = t1i & 3;
        int pv0 = __shfl_up_sync(FULL, nx0, 1), pv1 = __shfl_up_sync(FULL, nx1, 1);
        const int last0 = __shfl_sync(FULL, nx0, 31);
        if (lane == 0) {
            pv0 = 0;
            pv1 = last0;
        }
        bool ok = (t0i >> 2) == pv0;
        if (lane + 32 < 49) ok = ok && (t1i >> 2) == pv1;
        if (lane == 16) ok = ok && nx1 == 0;   // step 48 = the flush dibit
        if (__all_sync(FULL, ok)) {
            // dibit i = next state of step i; pack 4 dibits per byte
            unsigned v0 = (unsigned)nx0 << (6 - 2 * (lane & 3)), v1 = (lane + 32 < 48) ? (unsigned)nx1 << (6 - 2 * (lane & 3)) : 0u;
            v0 |= __shfl_xor_sync(FULL, v0, 1);
            v0 |= __shfl_xor_sync(FULL, v0, 2);
            v1 |= __shfl_xor_sync(FULL, v1, 1);
            v1 |= __shfl_xor_sync(FULL, v1, 2);
            if ((lane & 3) == 0) {
                out12[lane >> 2] = (unsigned char)v0;
                if (lane < 16) out12[8 + (lane >> 2)] = (unsigned char)v1;
            }
            __syncwarp();
            return 0;
        }
    }
    const int ns = (lane >> 2) & 3, ps = lane & 3;
    const int expect = T.trellis_pair[ps * 4 + ns];
    int m = ps == 0 ? 0 : (1 << 20);   // metric of state `ps` as seen by this lane
    for (int i = 0; i < 49; i++) {
        const int sym = __shfl_sync(FULL, i < 32 ? s0 : s1, i & 31);
        int key = ((m + __popc(expect ^ sym)) << 2) | ps;
        key = min(key, __shfl_xor_sync(FULL, key, 1));
        key = min(key, __shfl_xor_sync(FULL, key, 2));
        // key now holds the winner for next state `ns` in all four lanes of the group
        const unsigned b0 = __ballot_sync(FULL, key & 1), b1 = __ballot_sync(FULL, key & 2);
        if (lane == 0) {
            unsigned f = 0;
#pragma unroll
            for (int g = 0; g < 4; g++) f |= ((((b0 >> (4 * g)) & 1u) | (((b1 >> (4 * g)) & 1u) << 1)) << (2 * g));
            scratch[i] = (unsigned char)f;
        }
        m = __shfl_sync(FULL, key >> 2, ps * 4);   // new metric of state ps lives in group ps
    }
    const int m0 = __shfl_sync(FULL, m, 0);        // lane 0 has ps = 0
    __syncwarp();
    if (m0 > P25_VITERBI_MAX_FIX) return -1;
    if (lane == 0) {
        for (int i = 0; i < 12; i++) out12[i] = 0;
        int st = 0;
        for (int i = 48; i >= 0; i--) {
            if (i < 48) out12[i >> 2] |= (unsigned char)(st << (6 - 2 * (i & 3)));
            st = (scratch[i] >> (2 * st)) & 3;
        }
    }
    __syncwarp();
    return m0;
}

// 3/4-rate trellis (confirmed packet data), all 32 lanes: lane = (next state, two previous states pa and pa + 4);
// add-compare-select on the key (path metric << 3 | previous state) -- own pair, then two shuffle steps -- which keeps
// the lowest predecessor on ties exactly like p25_trellis_34_decode.  Survivors: 8 x 3 bits per step in surv[], traced
// back by lane 0.
__device__ __noinline__ int warp_trellis_34_decode(const P25DevTables& T, unsigned* surv, const unsigned char* dibits, int lane,
                                                   unsigned char* out18) {
    int s0 = 0, s1 = 0;
    {
        const int slot = T.interleave[lane];
        s0 = (dibits[2 * slot] << 2) | dibits[2 * slot + 1];
        if (lane + 32 < 49) {
            const int slot1 = T.interleave[lane + 32];
            s1 = (dibits[2 * slot1] << 2) | dibits[2 * slot1 + 1];
        }
    }
    const int ns = lane >> 2, pa = lane & 3, pb = pa + 4;
    const int ea = T.trellis34_pair[8 * pa + ns], eb = T.trellis34_pair[8 * pb + ns];
    int ma = pa == 0 ? 0 : (1 << 20), mb = 1 << 20;   // metrics of states pa and pb as seen by this lane
#pragma unroll 1
    for (int i = 0; i < 49; i++) {
        const int sym = __shfl_sync(FULL, i < 32 ? s0 : s1, i & 31);
        const int ka = ((ma + __popc(ea ^ sym)) << 3) | pa, kb = ((mb + __popc(eb ^ sym)) << 3) | pb;
        int key = min(ka, kb);
        key = min(key, __shfl_xor_sync(FULL, key, 1));
        key = min(key, __shfl_xor_sync(FULL, key, 2));
        const unsigned b0 = __ballot_sync(FULL, key & 1), b1 = __ballot_sync(FULL, key & 2), b2 = __ballot_sync(FULL, key & 4);
        if (lane == 0) {
            unsigned f = 0;
#pragma unroll
            for (int g = 0; g < 8; g++)
                f |= (((b0 >> (4 * g)) & 1u) | (((b1 >> (4 * g)) & 1u) << 1) | (((b2 >> (4 * g)) & 1u) << 2)) << (3 * g);
            surv[i] = f;
        }
        const int nm = key >> 3;                        // new metric of state ns, held by group ns
        ma = __shfl_sync(FULL, nm, 4 * pa);
        mb = __shfl_sync(FULL, nm, 4 * pb);
    }
    const int m0 = __shfl_sync(FULL, ma, 0);            // lane 0 has pa = 0
    __syncwarp();
    if (m0 > P25_VITERBI34_MAX_FIX) return -1;
    if (lane == 0) {
        for (int i = 0; i < P25_PDU_BLOCK34_BYTES; i++) out18[i] = 0;
        int st = 0;
        for (int i = 48; i >= 0; i--) {
            if (i < 48) {
                for (int b = 0; b < 3; b++) {
                    const int bit = 3 * i + b;
                    out18[bit >> 3] |= (unsigned char)(((st >> (2 - b)) & 1) << (7 - (bit & 7)));
                }
            }
            st = (surv[i] >> (3 * st)) & 7;
        }
    }
    __syncwarp();
    return m0;
}

// Berlekamp-Massey with a run-time syndrome count (one copy for RS(24,12), RS(24,16), RS(36,20)); same recurrence
// as p25_berlekamp_massey<N>.  Lane 0 only; all arrays in the warp's scratch area.
__device__ __noinline__ int bm_runtime(const P25DevTables& T, const unsigned char* S, unsigned char* lam, unsigned char* B,
                                       unsigned char* Tm, int N) {
    for (int i = 0; i <= N; i++) {
        lam[i] = 0;
        B[i] = 0;
    }
    lam[0] = 1;
    B[0] = 1;
    int L = 0, m = 1, b = 1;
#pragma unroll 1
    for (int r = 0; r < N; r++) {
        int d = S[r];
#pragma unroll 1
        for (int i = 1; i <= L; i++) d ^= p25_gf_mul(T, lam[i], S[r - i]);
        if (d == 0) {
            m++;
            continue;
        }
        const int coef = p25_gf_div(T, d, b);
        if (2 * L <= r) {
#pragma unroll 1
            for (int i = 0; i <= N; i++) Tm[i] = lam[i];
#pragma unroll 1
            for (int i = 0; i + m <= N; i++) lam[i + m] ^= (unsigned char)p25_gf_mul(T, coef, B[i]);
            L = r + 1 - L;
#pragma unroll 1
            for (int i = 0; i <= N; i++) B[i] = Tm[i];
            b = d;
            m = 1;
        } else {
#pragma unroll 1
            for (int i = 0; i + m <= N; i++) lam[i + m] ^= (unsigned char)p25_gf_mul(T, coef, B[i]);
            m++;
        }
    }
    return L;
}

__device__ __forceinline__ int warp_rs_syndromes(const P25DevTables& T, const unsigned char* sym, int n, int nroots, int lane) {
    int acc = 0;
    if (lane < nroots) {
        const int a = T.gf_exp[lane + 1];
#pragma unroll 1
        for (int i = 0; i < n; i++) acc = p25_gf_mul(T, acc, a) ^ sym[i];
    }
    return acc;
}

// Reed-Solomon (n, k) over GF(64), all 32 lanes: one syndrome per lane, Berlekamp-Massey on lane 0, Chien search and
// Forney magnitudes two positions per lane.  Same bounded-distance result (and the same "leave the word as received"
// on failure) as p25_rs_decode.  sym lives in shared memory; returns corrected symbols or -1, uniformly.
__device__ __noinline__ int warp_rs_decode(const P25DevTables& T, unsigned char* scr, unsigned char* sym, int n, int k, int lane) {
    const int nroots = n - k, t = nroots >> 1;
    unsigned char *S = scr, *lam = scr + 16, *omega = scr + 40, *dlam = scr + 56, *B = scr + 80, *Tm = scr + 104;
    {
        const int acc = warp_rs_syndromes(T, sym, n, nroots, lane);
        if (!__ballot_sync(FULL, acc != 0)) return 0;
        if (lane < nroots) S[lane] = (unsigned char)acc;
    }
    __syncwarp();
    int L = 0;
    if (lane == 0) L = bm_runtime(T, S, lam, B, Tm, nroots);
    L = __shfl_sync(FULL, L, 0);
    if (L > t) return -1;
    __syncwarp();
    if (lane < nroots) {
        int acc = 0;
#pragma unroll 1
        for (int j = 0; j <= lane && j <= L; j++) acc ^= p25_gf_mul(T, lam[j], S[lane - j]);
        omega[lane] = (unsigned char)acc;
    }
    if (lane < 17) dlam[lane] = (lane + 1 <= L && !(lane & 1)) ? lam[lane + 1] : 0;   // formal derivative: odd terms shift down
    __syncwarp();
    bool bad = false, isroot[2];
    int loc[2] = {0, 0}, mag[2] = {0, 0};
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int p = lane + 32 * h;
        const int xinv = T.gf_exp[(63 - p) % 63];
        isroot[h] = p < 63 && p25_poly_eval(T, lam, L, xinv) == 0;
        if (isroot[h]) {
            if (p >= n) {
                bad = true;
            } else {
                const int den = p25_poly_eval(T, dlam, L > 0 ? L - 1 : 0, xinv);
                const int mg = den ? p25_gf_div(T, p25_poly_eval(T, omega, nroots - 1, xinv), den) : 0;
                if (den == 0 || mg == 0) bad = true;
                loc[h] = n - 1 - p;
                mag[h] = mg;
            }
        }
    }
    const int roots = __popc(__ballot_sync(FULL, isroot[0])) + __popc(__ballot_sync(FULL, isroot[1]));
    if (__any_sync(FULL, bad) || roots != L) return -1;
#pragma unroll
    for (int h = 0; h < 2; h++)
        if (isroot[h]) sym[loc[h]] ^= (unsigned char)mag[h];
    __syncwarp();
    if (__ballot_sync(FULL, warp_rs_syndromes(T, sym, n, nroots, lane) != 0)) {
        __syncwarp();
#pragma unroll
        for (int h = 0; h < 2; h++)
            if (isroot[h]) sym[loc[h]] ^= (unsigned char)mag[h];   // leave the word as received
        __syncwarp();
        return -1;
    }
    return L;
}

// IMBE voice frame by the whole warp: four lanes gather each code word through the inverse interleave schedule,
// the PN mask bits come from the generator's closed form n steps ahead, and the seven code words are decoded side
// by side (lane group c owns c_c).  out15 (shared memory) receives u0..u7 and the 7 corrected-bit counts, exactly
// the output of p25_imbe_decode.
__device__ __noinline__ void warp_imbe_decode(const WalkShared& sh, const unsigned char* dibits, int lane, unsigned* out15) {
    const P25DevTables& T = sh.T;
    const int c = lane >> 2, q = lane & 3;
    const int nb = c < 7 ? T.imbe_cw_bits[c] : 7;
    unsigned v = 0;
    for (int b = q; b < nb; b += 4) {
        const int i = sh.imbe_src[c * 24 + b];
        v |= ((dibits[i >> 1] >> (1 - (i & 1))) & 1u) << b;
    }
    v |= __shfl_xor_sync(FULL, v, 1);
    v |= __shfl_xor_sync(FULL, v, 2);
    unsigned u0 = 0;
    const int e0 = p25_golay23_decode(T, v, &u0);   // meaningful on group 0
    u0 = __shfl_sync(FULL, u0, 0);
    const unsigned p0 = (16u * u0) & 0xFFFF;
    // PN bits of code word c start after those of code words 1 .. c-1 (23, 23, 23, 15, 15, 15 bits)
    const int off = c <= 4 ? 23 * (c - 1) : 69 + 15 * (c - 4);
    unsigned mask = 0;
    if (c >= 1 && c <= 6) {
        for (int j = q; j < nb; j += 4) {
            const int n = off + 1 + j;
            const unsigned pn = (sh.pn_a[n] * p0 + sh.pn_c[n]) & 0xFFFF;
            mask |= (pn >> 15) << (nb - 1 - j);
        }
    }
    mask |= __shfl_xor_sync(FULL, mask, 1);
    mask |= __shfl_xor_sync(FULL, mask, 2);
    const unsigned w = v ^ mask;
    unsigned dg = 0, dh = 0;
    const int eg = p25_golay23_decode(T, w, &dg), eh = p25_hamming15_decode(T, w, &dh);
    if (q == 0) {
        if (c == 0) {
            out15[0] = u0;
            out15[8] = (unsigned)e0;
        } else if (c < 4) {
            out15[c] = dg;
            out15[8 + c] = (unsigned)eg;
        } else if (c < 7) {
            out15[c] = dh;
            out15[8 + c] = (unsigned)eh;
        } else {
            out15[7] = v & 0x7F;
        }
    }
    __syncwarp();
}

// TSBK block decoded by the warp (reference src/recv.rs:231); lane 0 applies the result.
__device__ __forceinline__ void tsbk_apply(const WarpCtx& c, unsigned long long idx, int fixed, const unsigned char* out) {
    WalkState& ws = *c.ws;
    ws.cnt = 0;
    if (fixed < 0) {
        stat_bad(c, P25CU_ST_VITERBI_DIBIT);
        fail(c, P25CU_E_VITERBI_DIBIT, idx);
        return;
    }
    stat_ok(c, P25CU_ST_VITERBI_DIBIT, (unsigned)fixed);
    ws.blocks++;
    if ((out[0] & 0x80) || ws.blocks == 3) ws.state = WS_FLUSH;
    emit(c, P25CU_EV_TSBK, idx, out, 12);
}

// hexbits (6 bits each, MSB first) -> bytes, one output byte per lane (same packing as p25_pack_hexbits)
__device__ __forceinline__ void warp_pack_hexbits(const unsigned char* hex, int nhex, unsigned char* out, int lane) {
    if (lane < nhex * 6 / 8) {
        unsigned v = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int b = 8 * lane + k;
            v = (v << 1) | ((hex[b / 6] >> (5 - b % 6)) & 1u);
        }
        out[lane] = (unsigned char)v;
    }
    __syncwarp();
}

// lane 0 adds a batch of code-word outcomes to a stats family (n words, `bad` of them uncorrectable, `fixed` bits)
__device__ __forceinline__ void stat_batch(const WarpCtx& c, int fam, unsigned n, unsigned bad, unsigned fixed) {
    unsigned long long* s = c.p->stats + ((size_t)c.stream * P25CU_ST_FAMILIES + fam) * 3;
    s[0] += n;
    s[1] += bad;
    s[2] += fixed;
}

// A payload decision point has been reached (ws.cnt == target).  Called by the whole warp: code words of a batch
// are decoded one per lane, Reed-Solomon and IMBE frames by the warp-cooperative decoders above; lane 0 alone
// touches the receiver state, the stats and the event slots.
__device__ __noinline__ void complete_payload(const WarpCtx& c, const WalkShared& sh, unsigned char* scr, unsigned long long idx, int lane) {
    WalkState& ws = *c.ws;
    const P25DevTables& T = *c.T;
    const int duid = ws.duid;
    if (duid == 0x0 || duid == 0xF) {
        // HDU -> VoiceHeader (reference src/recv.rs:223): 36 x Golay(18,6) + RS(36,20)
        // TDULC -> VoiceTerm (reference src/recv.rs:232): 12 x Golay(24,12) + RS(24,12)
        const bool hdu = duid == 0x0;
        const int nw = hdu ? 36 : 12;
        unsigned nbad = 0, nfix = 0;
        for (int w = lane; w < nw; w += 32) {
            if (hdu) {
                unsigned d6;
                const int n = p25_golay18_decode(T, p25_take_bits(ws.buf, 18 * w, 18), &d6);
                if (n < 0) nbad++; else nfix += (unsigned)n;
                ws.hex[w] = (unsigned char)d6;
            } else {
                unsigned d12;
                const unsigned word = p25_take_bits(ws.buf, 24 * w, 24);
                const int n = p25_golay24_decode(T, word, &d12);
                if (n < 0) {
                    nbad++;
                    d12 = word >> 12;
                } else {
                    nfix += (unsigned)n;
                }
                ws.hex[2 * w] = (unsigned char)(d12 >> 6);
                ws.hex[2 * w + 1] = (unsigned char)(d12 & 0x3F);
            }
        }
        nbad = __reduce_add_sync(FULL, nbad);
        nfix = __reduce_add_sync(FULL, nfix);
        __syncwarp();
        const int n = warp_rs_decode(T, scr, ws.hex, hdu ? 36 : 24, hdu ? 20 : 12, lane);
        __syncwarp();
        if (n >= 0) warp_pack_hexbits(ws.hex, hdu ? 20 : 12, c.payload, lane);
        if (lane == 0) {
            stat_batch(c, hdu ? P25CU_ST_GOLAY_SHORT : P25CU_ST_GOLAY_EXT, (unsigned)nw, nbad, nfix);
            const int fam = hdu ? P25CU_ST_RS_LONG : P25CU_ST_RS_SHORT;
            if (n < 0) {
                stat_bad(c, fam);
                fail(c, P25CU_E_RS, idx);
            } else {
                stat_ok(c, fam, (unsigned)n);
                ws.state = WS_FLUSH;
                emit(c, hdu ? P25CU_EV_VOICE_HEADER : P25CU_EV_VOICE_TERM, idx, c.payload, hdu ? 15 : 9);
            }
        }
        return;
    }
    // LDU1 / LDU2 parts (reference src/recv.rs:224-230)
    const int part = ws.part;
    const int kind = T.ldu_kind[part];
    const unsigned char* d = ws.buf + T.ldu_start[part];
    if (kind == 0) {
        unsigned* pl = reinterpret_cast<unsigned*>(scr + 128);   // 15 words
        warp_imbe_decode(sh, d, lane, pl);
        if (lane == 0) {
            ws.part = part + 1;
            stat_batch(c, P25CU_ST_GOLAY_STD, 4, 0, pl[8] + pl[9] + pl[10] + pl[11]);
            stat_batch(c, P25CU_ST_HAMMING_STD, 3, 0, pl[12] + pl[13] + pl[14]);
            if (part + 1 == P25_LDU_PARTS) ws.state = WS_FLUSH;
            emit(c, P25CU_EV_VOICE_FRAME, idx, pl, 60);
        }
        return;
    }
    if (kind == 1) {
        const int chunks = ws.chunks;
        unsigned bad = 0, fix = 0;
        if (lane < 4) {
            unsigned d6;
            const int n = p25_hamming10_decode(T, p25_take_bits(d, 10 * lane, 10), &d6);
            if (n < 0) bad = 1; else fix = (unsigned)n;
            ws.hex[4 * chunks + lane] = (unsigned char)d6;
        }
        bad = __reduce_add_sync(FULL, bad);
        fix = __reduce_add_sync(FULL, fix);
        __syncwarp();
        const bool lc = duid == 0x5;
        int n = 0;
        if (chunks + 1 == 6) {
            n = warp_rs_decode(T, scr, ws.hex, 24, lc ? 12 : 16, lane);
            __syncwarp();
            if (n >= 0) warp_pack_hexbits(ws.hex, lc ? 12 : 16, c.payload, lane);
        }
        if (lane == 0) {
            ws.part = part + 1;
            ws.chunks = chunks + 1;
            stat_batch(c, P25CU_ST_HAMMING_SHORT, 4, bad, fix);
            if (chunks + 1 == 6) {
                const int fam = lc ? P25CU_ST_RS_SHORT : P25CU_ST_RS_MED;
                if (n < 0) {
                    stat_bad(c, fam);
                    fail(c, P25CU_E_RS, idx);
                } else {
                    stat_ok(c, fam, (unsigned)n);
                    emit(c, lc ? P25CU_EV_LINK_CONTROL : P25CU_EV_CRYPTO_CONTROL, idx, c.payload, lc ? 9 : 12);
                }
            }
        }
        return;
    }
    {   // low-speed data: 2 x cyclic(16,8)
        unsigned d8 = 0, bad = 0, fix = 0;
        if (lane < 2) {
            const int n = p25_cyclic16_decode(T, p25_take_bits(d, 16 * lane, 16), &d8);
            if (n < 0) bad = 1; else fix = (unsigned)n;
        }
        bad = __reduce_add_sync(FULL, bad);
        fix = __reduce_add_sync(FULL, fix);
        const unsigned d1 = __shfl_sync(FULL, d8, 1);
        if (lane == 0) {
            ws.part = part + 1;
            stat_batch(c, P25CU_ST_CYCLIC, 2, bad, fix);
            const unsigned frag = (d8 << 8) | d1;
            emit(c, P25CU_EV_LSD, idx, &frag, 4);
        }
    }
}

// One 98-dibit block of a packet data unit (DUID 0xC).  The header block is 1/2-rate coded and announces the number of
// data blocks and their format; confirmed data blocks are 3/4-rate coded, unconfirmed ones 1/2-rate [STD].  No
// MessageEvent variant carries packet data (reference src/recv.rs:214-233): only the viterbiDibit / viterbiTribit
// stats families (src/hub.rs:569-570) and P25Error values leave the receiver.  ws.blocks = blocks decoded so far,
// ws.part = data blocks announced, ws.chunks = confirmed format.  Called by the whole warp.
__device__ __noinline__ void pdu_block(const WarpCtx& c, WalkShared& sh, int warp, unsigned long long idx, int lane) {
    WalkState& ws = *c.ws;
    unsigned char* out = c.payload;                      // 18 bytes at most
    const bool header = ws.blocks == 0;
    const bool tri = !header && ws.chunks;
    int fixed;
    if (tri) fixed = warp_trellis_34_decode(sh.T, sh.surv[warp], ws.buf, lane, out);
    else fixed = warp_trellis_half_decode(sh.T, sh.scr[warp], ws.buf, lane, out);
    __syncwarp();
    if (lane == 0) {
        ws.cnt = 0;
        const int fam = tri ? P25CU_ST_VITERBI_TRIBIT : P25CU_ST_VITERBI_DIBIT;
        if (fixed < 0) {
            stat_bad(c, fam);
            fail(c, tri ? P25CU_E_VITERBI_TRIBIT : P25CU_E_VITERBI_DIBIT, idx);
        } else {
            stat_ok(c, fam, (unsigned)fixed);
            if (header) {
                if (p25_crc_ccitt(out, 10) != (((unsigned)out[10] << 8) | out[11])) {
                    enter_sync(ws, idx + 1);             // the length field cannot be trusted: drop lock
                } else {
                    ws.blocks = 1;
                    ws.part = out[6] & 0x7F;
                    ws.chunks = (out[0] & 0x1F) == P25_PDU_FORMAT_CONFIRMED;
                    if (ws.part == 0) ws.state = WS_FLUSH;
                }
            } else if (++ws.blocks == 1 + ws.part) {
                ws.state = WS_FLUSH;
            }
        }
    }
}

// ---------------------------------------------------------------- the walker
// Stage the decode tables and build the derived ones (callers __syncthreads() before use).
__device__ __forceinline__ void walk_shared_init(WalkShared& sh, const P25DevTables* tables) {
    const uint4* src = (const uint4*)tables;
    uint4* dst = (uint4*)&sh.T;
    for (unsigned i = threadIdx.x; i < sizeof(P25DevTables) / 16; i += blockDim.x) dst[i] = src[i];
    if (threadIdx.x == 0) {   // IMBE PN generator n steps ahead: p_n = a_n * p_0 + c_n (mod 2^16), a_n = 173^n
        unsigned a = 1, cc = 0;
        for (int n = 0; n < 116; n++) {
            sh.pn_a[n] = (unsigned short)a;
            sh.pn_c[n] = (unsigned short)cc;
            a = (173u * a) & 0xFFFF;
            cc = (173u * cc + 13849u) & 0xFFFF;
        }
    }
    for (unsigned i = threadIdx.x; i < 144; i += blockDim.x) sh.imbe_src[tables->imbe_cw[i] * 24 + tables->imbe_bit[i]] = (unsigned char)i;

}
// PRE: consult the tensor-pipe prefilter before the exact correlator (contexts whose streams are mostly idle: the
// channelizer's 1,536 slots per capture).  Decoded output is identical either way; where every stream carries a signal
// the extra code in the search loop only costs (cfg5: +6 % walker time, A/B in profiles/r02_walk_prefilter_ab.txt).
template <bool PRE>
__global__ void __launch_bounds__(32 * P25CU_WALK_WARPS, 16) p25_walk_kernel(const WalkParams p) {
    __shared__ WalkShared sh;
    walk_shared_init(sh, p.tables);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // The prefilter's operands exist only in the instantiation that uses it: beside the /50 demod kernel (3 x 68,096 bytes
    // of dynamic shared memory + 1 KB per CTA of system use, 228 KB per SM) two walker CTAs have 12,032 bytes each --
    // p25_walk_kernel<false> must stay below that or the co-running pair no longer fits one SM.
    static_assert(sizeof(WalkShared) <= 12032, "p25_walk_kernel<false> has to fit beside three /50 demod CTAs, twice");
    const unsigned* fpb = nullptr;
    float* pe = nullptr;
    if constexpr (PRE) {
        __shared__ unsigned s_fpb[FPB_LEN];          // sync fingerprint as BF16 pairs: (fp[i - 8], fp[i - 7]), zeros outside
        __shared__ float s_pe[P25CU_WALK_WARPS][48];   // prefix sums of the 8-sample block energies of the staged window
        for (int i = threadIdx.x; i < FPB_LEN; i += blockDim.x) {
            const float lo = (i >= 8 && i < 8 + P25_FP_LEN) ? c_sync_fp[i - 8] : 0.f, hi = (i >= 7 && i < 7 + P25_FP_LEN) ? c_sync_fp[i - 7] : 0.f;
            s_fpb[i] = pack_bf16(lo, hi);
        }
        fpb = s_fpb;
        pe = s_pe[warp];
    }
    WalkState& ws = sh.ws[warp];
    __syncthreads();                                          // tables staged
    // grid-stride over the streams: one pass when the grid covers them all (the usual launch), several when the walker
    // is launched as a small persistent grid that fits beside the next chunk's demod CTAs (p25cu_launch_walk)
  for (unsigned stream = blockIdx.x * P25CU_WALK_WARPS + warp; stream < p.n_streams; stream += gridDim.x * P25CU_WALK_WARPS) {
    __syncwarp();
    {   // header, then the dibits of the unit in flight (kept 4 per byte in HBM, one per byte here)
        const uint4* src = (const uint4*)(p.states + stream);
        uint4* dst = (uint4*)&ws;
        if (lane < (int)(sizeof(WalkHeader) / 16)) dst[lane] = src[lane];
        __syncwarp();
        const int st0 = ws.state;
        const int held = (st0 == WS_NID || st0 == WS_PAYLOAD) ? ws.cnt : 0;
        const unsigned* pk = p.states[stream].packed;
        for (int j = lane; 16 * j < held; j += 32) {
            const unsigned w = pk[j];
            unsigned o[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const unsigned b = (w >> (8 * q)) & 0xFFu;
                o[q] = (b & 3u) | (((b >> 2) & 3u) << 8) | (((b >> 4) & 3u) << 16) | (((b >> 6) & 3u) << 24);
            }
            reinterpret_cast<uint4*>(ws.buf)[j] = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
    WarpCtx c{&p, &sh.T, &ws, stream, &sh.pend[warp], sh.scr[warp] + 128};
    if (lane == 0) sh.pend[warp].valid = 0;
    __syncwarp();
    float* win = sh.win[warp];
    const unsigned long long p0 = p.p0, end = p.p0 + p.n;
    const float* row = p.bb + (size_t)stream * p.row_stride;  // row[i] <-> absolute sample p0 - 256 + i
    const unsigned lane_le = 0xFFFFFFFFu >> (31 - lane);

    if (ws.resync_req) {  // MessageReceiver::resync, reference src/recv.rs:136
        __syncwarp();
        if (lane == 0) {
            enter_sync(ws, p0);
            ws.resync_req = 0;
        }
        __syncwarp();
    }

    for (;;) {
        const int st = ws.state;
        if (st == WS_SYNC) {
            // ---- frame-sync search over 128 positions: lane l owns the four consecutive positions pos + 4 l + h.
            // Position q correlates win[q .. q + 230]; a lane's four windows overlap in all but three samples, so the
            // samples of four taps come from two 16-byte loads (the window is kept twice, the second copy shifted by
            // one sample, which makes every packed operand an aligned register pair) and the second half of each
            // load is reused by the next tap group: 2 LDS.128 per 16 FFMA2 instead of 16 LDS.32.  Each position's sum
            // still runs over k = 0 .. 230 in order with one fmaf per tap, exactly like the oracle.
            unsigned long long pos = ws.pos;
            if (pos >= end) break;
            const long long wlim = (long long)P25CU_BB_HIST + (long long)p.n;                   // first invalid row index
            float* win1 = win + WIN_PAD;
            // An idle or noise-only channel stays in this loop -- stage one
            // copy of the window, fifteen tensor steps, next 128 positions -- and never runs the exact correlator again;
            // a step without a candidate leaves the detector's carried state at "previous not above" (prev_corr is only
            // ever compared when the previous position was above).
            if (PRE && (int)ws.quiet >= P25_PREFILTER_QUIET) {
                const unsigned long long pos_in = pos;
                bool hit = false;
                while (pos < end) {
                    const long long wb = (long long)(pos - p0) + P25CU_BB_HIST - (P25_FP_LEN - 1);
                    // start the staged window on a 16-byte boundary of the row (three LDG.128 per lane instead of twelve
                    // LDG.32): the tile then begins up to three positions early -- testing a position twice is harmless
                    const int o = (int)(wb & 3);
                    const bool vec = wb - o + WIN_PAD <= wlim;
                    __syncwarp();
                    if (vec) {
                        const float4* src = reinterpret_cast<const float4*>(row + (wb - o));
                        for (int i = lane; i < WIN_PAD / 4; i += 32) reinterpret_cast<float4*>(win)[i] = __ldg(src + i);
                        // the next step's 128 new samples: in flight while this step computes (the row comes from HBM)
                        if (lane < 5 && wb - o + WIN_PAD + 32 * lane < wlim) prefetch_l1(row + (wb - o) + WIN_PAD + 32 * lane);
                    } else {
                        for (int i = lane; i < WIN_PAD; i += 32) win[i] = (i < WIN_LEN && wb + i < wlim) ? __ldg(row + wb + i) : 0.f;
                    }
                    __syncwarp();
                    hit = sync_prefilter(win, fpb, pe, lane);
                    if (hit) break;
                    const unsigned long long left0 = end - pos, step = (unsigned long long)(SEARCH_N - (vec ? o : 0));
                    pos += left0 < step ? left0 : step;
                }
                if (pos != pos_in) {
                    __syncwarp();
                    if (lane == 0) {
                        ws.prev_above = 0;
                        ws.have_prev = 1;
                        ws.pos = pos;
                    }
                    __syncwarp();
                }
                if (!hit) break;                                   // the chunk is exhausted
            }
            {   // a candidate (or a stream that is not idle): both copies of the window at pos for the exact correlator
                const long long wbase = (long long)(pos - p0) + P25CU_BB_HIST - (P25_FP_LEN - 1);   // row index of win[0]
                __syncwarp();
                for (int i = lane; i < WIN_PAD; i += 32) {
                    const float v = (i < WIN_LEN && wbase + i < wlim) ? __ldg(row + wbase + i) : 0.f;
                    win[i] = v;
                    if (i) win1[i - 1] = v;
                }
                if (lane < 5 && wbase + WIN_LEN + 32 * lane < wlim) prefetch_l1(row + wbase + WIN_LEN + 32 * lane);   // next step's new samples
                __syncwarp();
            }
            float2 c01 = make_float2(0.f, 0.f), c23 = c01, e01 = c01, e23 = c01;   // positions h = 0,1 | 2,3
            {
                const float4* wa = reinterpret_cast<const float4*>(win) + lane;    // wa[m] = win[4 (lane + m) ..]
                const float4* wb = reinterpret_cast<const float4*>(win1) + lane;   // wb[m] = win[4 (lane + m) + 1 ..]
                float4 a = wa[0], b = wb[0];
#pragma unroll 2
                for (int m = 0; m < P25_FP_LEN / 4; m++) {                         // 57 full groups of 4 taps
                    const float4 an = wa[m + 1], bn = wb[m + 1];
                    const float f0 = c_sync_fp[4 * m], f1 = c_sync_fp[4 * m + 1], f2 = c_sync_fp[4 * m + 2], f3 = c_sync_fp[4 * m + 3];
                    float2 x01 = make_float2(a.x, a.y), x23 = make_float2(a.z, a.w);          // tap 4m:     win[q + 4m]
                    c01 = fma2(make_float2(f0, f0), x01, c01);
                    c23 = fma2(make_float2(f0, f0), x23, c23);
                    e01 = fma2(x01, x01, e01);
                    e23 = fma2(x23, x23, e23);
                    x01 = make_float2(b.x, b.y), x23 = make_float2(b.z, b.w);                 // tap 4m + 1
                    c01 = fma2(make_float2(f1, f1), x01, c01);
                    c23 = fma2(make_float2(f1, f1), x23, c23);
                    e01 = fma2(x01, x01, e01);
                    e23 = fma2(x23, x23, e23);
                    x01 = make_float2(a.z, a.w), x23 = make_float2(an.x, an.y);               // tap 4m + 2
                    c01 = fma2(make_float2(f2, f2), x01, c01);
                    c23 = fma2(make_float2(f2, f2), x23, c23);
                    e01 = fma2(x01, x01, e01);
                    e23 = fma2(x23, x23, e23);
                    x01 = make_float2(b.z, b.w), x23 = make_float2(bn.x, bn.y);               // tap 4m + 3
                    c01 = fma2(make_float2(f3, f3), x01, c01);
                    c23 = fma2(make_float2(f3, f3), x23, c23);
                    e01 = fma2(x01, x01, e01);
                    e23 = fma2(x23, x23, e23);
                    a = an;
                    b = bn;
                }
                {   // taps 228, 229, 230
                    const float4 an = wa[P25_FP_LEN / 4 + 1];
                    const float f0 = c_sync_fp[P25_FP_LEN - 3], f1 = c_sync_fp[P25_FP_LEN - 2], f2 = c_sync_fp[P25_FP_LEN - 1];
                    float2 x01 = make_float2(a.x, a.y), x23 = make_float2(a.z, a.w);
                    c01 = fma2(make_float2(f0, f0), x01, c01);
                    c23 = fma2(make_float2(f0, f0), x23, c23);
                    e01 = fma2(x01, x01, e01);
                    e23 = fma2(x23, x23, e23);
                    x01 = make_float2(b.x, b.y), x23 = make_float2(b.z, b.w);
                    c01 = fma2(make_float2(f1, f1), x01, c01);
                    c23 = fma2(make_float2(f1, f1), x23, c23);
                    e01 = fma2(x01, x01, e01);
                    e23 = fma2(x23, x23, e23);
                    x01 = make_float2(a.z, a.w), x23 = make_float2(an.x, an.y);
                    c01 = fma2(make_float2(f2, f2), x01, c01);
                    c23 = fma2(make_float2(f2, f2), x23, c23);
                    e01 = fma2(x01, x01, e01);
                    e23 = fma2(x23, x23, e23);
                }
            }
            const unsigned long long left = end - pos;
            const int nvalid = left < SEARCH_N ? (int)left : SEARCH_N;
            const float corr[4] = {c01.x, c01.y, c23.x, c23.y}, en[4] = {e01.x, e01.y, e23.x, e23.y};
            bool ab[4];
#pragma unroll
            for (int h = 0; h < 4; h++)
                ab[h] = 4 * lane + h < nvalid && corr[h] > 0.f && (corr[h] * corr[h] >= P25_SYNC_RHO2_EFP * en[h]);
            // the detector fires on the first position that is above threshold, whose predecessor was above too, and
            // whose correlation did not grow; the predecessor of a lane's first position is the lane below's last one
            float pc = __shfl_up_sync(FULL, corr[3], 1);
            int pa = __shfl_up_sync(FULL, (int)ab[3], 1), hp = 1;
            if (lane == 0) {
                pc = ws.prev_corr;
                pa = ws.prev_above;
                hp = ws.have_prev;
            }
            const bool any_above = PRE && __any_sync(FULL, ab[0] || ab[1] || ab[2] || ab[3]);
            unsigned cand = 0xFFFFFFFFu;
#pragma unroll
            for (int h = 0; h < 4; h++) {
                if (cand == 0xFFFFFFFFu && hp && ab[h] && pa && corr[h] <= pc) cand = 4 * lane + h;
                pc = corr[h];
                pa = (int)ab[h];
                hp = 1;
            }
            const unsigned first = __reduce_min_sync(FULL, cand);
            const int fire_off = first == 0xFFFFFFFFu ? -1 : (int)first;
            // state handed to the next step: the last valid position of this one
            const int ql = nvalid - 1, hl = ql & 3;
            const float csel = hl == 0 ? corr[0] : hl == 1 ? corr[1] : hl == 2 ? corr[2] : corr[3];
            const int asel = (int)(hl == 0 ? ab[0] : hl == 1 ? ab[1] : hl == 2 ? ab[2] : ab[3]);
            const float lc = __shfl_sync(FULL, csel, ql >> 2);
            const int la = __shfl_sync(FULL, asel, ql >> 2);
            if (fire_off < 0) {
                __syncwarp();
                if (lane == 0) {
                    ws.prev_corr = lc;
                    ws.prev_above = la;
                    ws.have_prev = 1;
                    ws.pos = pos + nvalid;
                    if (PRE) ws.quiet = any_above ? 0 : (ws.quiet < 255 ? ws.quiet + 1 : 255);
                }
                __syncwarp();
                continue;
            }
            // ---- lock: the correlation peaked on the sample before the firing one
            const unsigned long long idx = pos + fire_off;
            const long long pk = (long long)idx - 1;
            float v = 0.f;
            if (lane < P25_FS_DIBITS) v = row[(pk - (long long)p0) + P25CU_BB_HIST - (P25_FP_LEN - 1) + P25_SPS * lane];
            float ps = 0.f, ns = 0.f;
#pragma unroll
            for (int i = 0; i < P25_FS_DIBITS; i++) {
                const float vi = __shfl_sync(FULL, v, i);
                if ((P25_SYNC_POS_MASK >> i) & 1)
                    ps += vi;
                else
                    ns += vi;
            }
            const float pavg = ps / 11.0f, navg = ns / 13.0f;
            const float mid = (pavg + navg) * 0.5f;
            const float pth = mid + (pavg - mid) * (2.0f / 3.0f);
            const float nth = mid + (navg - mid) * (2.0f / 3.0f);
            __syncwarp();
            if (lane == 0) {
                ws.mid = mid;
                ws.pth = pth;
                ws.nth = nth;
                ws.next_sym = (unsigned long long)(pk + P25_SPS);
                ws.frame_pos = P25_FS_DIBITS;
                ws.state = WS_NID;
                ws.cnt = 0;
                if (PRE) ws.quiet = 0;
            }
            __syncwarp();
            continue;
        }

        // ---- symbol states: slice up to 32 symbol instants
        const unsigned long long t0 = ws.next_sym;
        if (t0 >= end) break;
        const int cnt = ws.cnt;
        const unsigned fpos0 = ws.frame_pos;
        int target;
        if (st == WS_NID) {
            target = P25_NID_DIBITS;
        } else if (st == WS_PAYLOAD) {
            const int duid = ws.duid;
            if (duid == 0x7 || duid == 0xC) target = P25_TSBK_DIBITS;
            else if (duid == 0x0) target = P25_HDU_DIBITS;
            else if (duid == 0xF) target = P25_TDULC_DIBITS;
            else target = sh.T.ldu_start[ws.part] + sh.T.ldu_len[ws.part];
        } else {
            target = 0x7FFFFFFF;
        }
        const int need = target - cnt;
        // 64 symbol instants per step: lane l slices instants l and 32 + l
        const unsigned long long span = (end - 1 - t0) / P25_SPS + 1;
        const int avail = span < 64 ? (int)span : 64;
        const bool va = lane < avail, vb = lane + 32 < avail;
        const bool sta = ((fpos0 + lane) % P25_STATUS_PERIOD) == P25_STATUS_PERIOD - 1;
        const bool stb = ((fpos0 + 32 + lane) % P25_STATUS_PERIOD) == P25_STATUS_PERIOD - 1;
        const float* sp = row + (long long)(t0 - p0) + P25CU_BB_HIST + P25_SPS * lane;
        const float sa = va ? __ldg(sp) : 0.f, sb = vb ? __ldg(sp + 32 * P25_SPS) : 0.f;
        {   // the 64 symbol instants of the next step span 640 samples = 20 cache lines
            const long long nx = (long long)(t0 - p0) + P25CU_BB_HIST + 64 * P25_SPS + 32 * lane;
            if (lane < 21 && nx < (long long)P25CU_BB_HIST + (long long)p.n) prefetch_l1(row + nx);
        }
        const float pth = ws.pth, mid = ws.mid, nth = ws.nth;
        const int da = sa > pth ? 1 : (sa > mid ? 0 : (sa > nth ? 2 : 3));   // [STD] 01 +3, 00 +1, 10 -1, 11 -3
        const int db = sb > pth ? 1 : (sb > mid ? 0 : (sb > nth ? 2 : 3));
        const unsigned dm_a = __ballot_sync(FULL, va && !sta), dm_b = __ballot_sync(FULL, vb && !stb);
        const int ca = __popc(dm_a & lane_le), cb = __popc(dm_a) + __popc(dm_b & lane_le);
        unsigned st_a, st_b;
        if (st == WS_FLUSH) {
            st_a = __ballot_sync(FULL, va && sta);
            st_b = __ballot_sync(FULL, vb && stb);
        } else {
            st_a = __ballot_sync(FULL, va && !sta && ca == need);
            st_b = __ballot_sync(FULL, vb && !stb && cb == need);
        }
        const unsigned stopmask = st_a | st_b;
        const int K = st_a ? __ffs(st_a) : (st_b ? 32 + __ffs(st_b) : avail);
        if (st != WS_FLUSH) {
            if (lane < K && va && !sta) ws.buf[cnt + ca - 1] = (unsigned char)da;
            if (lane + 32 < K && vb && !stb) ws.buf[cnt + cb - 1] = (unsigned char)db;
        }
        const unsigned mk_a = K >= 32 ? FULL : ((1u << K) - 1u);
        const unsigned mk_b = K <= 32 ? 0u : (K >= 64 ? FULL : ((1u << (K - 32)) - 1u));
        const int ndata = __popc(dm_a & mk_a) + __popc(dm_b & mk_b);
        __syncwarp();
        if (lane == 0) {
            ws.next_sym = t0 + (unsigned long long)P25_SPS * K;
            ws.frame_pos = fpos0 + K;
            if (st != WS_FLUSH) ws.cnt = cnt + ndata;
        }
        __syncwarp();
        if (stopmask) {
            const unsigned long long idx = t0 + (unsigned long long)P25_SPS * (K - 1);
            if (st == WS_FLUSH) {
                if (lane == 0) enter_sync(ws, idx + 1);
            } else if (st == WS_NID) {
                // the 64 NID bits, dibit i at bits 63 - 2 i .. 62 - 2 i: one dibit per lane, two warp OR-reductions
                static_assert(P25_NID_DIBITS == 32, "one NID dibit per lane");
                const unsigned long long mine = (unsigned long long)(ws.buf[lane] & 3u) << (2 * (31 - lane));
                const unsigned long long bits = ((unsigned long long)__reduce_or_sync(FULL, (unsigned)(mine >> 32)) << 32) |
                                                __reduce_or_sync(FULL, (unsigned)mine);
                unsigned data = 0;
                const int nerr = warp_bch_decode(sh.T, sh.scr[warp], bits >> 1, lane, &data);
                if (lane == 0) nid_apply(c, idx, nerr, data);
            } else if (ws.duid == 0x7) {
                unsigned char* out = ws.hex;   // 12-byte result area
                const int fixed = warp_trellis_half_decode(sh.T, sh.scr[warp], ws.buf, lane, out);
                if (lane == 0) tsbk_apply(c, idx, fixed, out);
            } else if (ws.duid == 0xC) {
                pdu_block(c, sh, warp, idx, lane);
            } else {
                complete_payload(c, sh, sh.scr[warp], idx, lane);
            }
            __syncwarp();
            flush_event(c, lane);
        }
    }

    // ---- end of chunk: persist state, hand the last 256 samples to the front of the next chunk's row
    __syncwarp();
    {
        uint4* dst = (uint4*)(p.states + stream);
        const uint4* src = (const uint4*)&ws;
        if (lane < (int)(sizeof(WalkHeader) / 16)) dst[lane] = src[lane];
        const int st1 = ws.state;
        const int held = (st1 == WS_NID || st1 == WS_PAYLOAD) ? ws.cnt : 0;
        unsigned* pk = p.states[stream].packed;
        for (int j = lane; 16 * j < held; j += 32) {
            const uint4 v = reinterpret_cast<const uint4*>(ws.buf)[j];
            const unsigned in[4] = {v.x, v.y, v.z, v.w};
            unsigned w = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const unsigned x = in[q];
                w |= ((x & 3u) | (((x >> 8) & 3u) << 2) | (((x >> 16) & 3u) << 4) | (((x >> 24) & 3u) << 6)) << (8 * q);
            }
            pk[j] = w;
        }
    }
    float* nxt = p.bb_next + (size_t)stream * p.row_stride;
    float keep[P25CU_BB_HIST / 32];
#pragma unroll
    for (int i = 0; i < P25CU_BB_HIST / 32; i++) keep[i] = row[p.n + lane + 32 * i];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < P25CU_BB_HIST / 32; i++) nxt[lane + 32 * i] = keep[i];
  }
}

cudaError_t p25cu_walk_upload_consts() {
    return cudaMemcpyToSymbol(c_sync_fp, P25_SYNC_FP, sizeof(float) * P25_FP_LEN);
}

cudaError_t p25cu_launch_walk(const WalkParams& p, cudaStream_t st, unsigned max_blocks, int device, bool prefilter) {
    // Beside a demod kernel (max_blocks != 0) keep the SM's shared-memory carve-out at its maximum: that kernel needs
    // ~204 KB per SM and an SM cannot change its carve-out while any CTA is resident on it -- a walker CTA that asked
    // for the default (L1-heavy) split kept the demod CTAs off its SM until it exited (measured: 0.48 -> 0.71 ms).
    // Alone, the walker prefers the L1-heavy default (its row reads hit L1; 0.089 vs 0.104 ms).
    // The attribute is a per-device property of the function and contexts on several devices (or two contexts on one
    // device with different overlap settings) may launch from different host threads: the last value set is tracked
    // per device and the set + launch pair runs under a mutex.
    static std::mutex mu;
    static int carve_state[2][P25CU_MAX_DEVICES];
    static bool init = false;
    std::lock_guard<std::mutex> lk(mu);
    if (!init) {
        for (int i = 0; i < P25CU_MAX_DEVICES; i++) carve_state[0][i] = carve_state[1][i] = -2;
        init = true;
    }
    const int want = max_blocks ? (int)cudaSharedmemCarveoutMaxShared : (int)cudaSharedmemCarveoutDefault;
    if (device < 0 || device >= P25CU_MAX_DEVICES) return cudaErrorInvalidDevice;
    if (carve_state[prefilter][device] != want) {
        cudaError_t e = prefilter ? cudaFuncSetAttribute(p25_walk_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, want)
                                  : cudaFuncSetAttribute(p25_walk_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, want);
        if (e != cudaSuccess) return e;
        carve_state[prefilter][device] = want;
    }
    unsigned blocks = (p.n_streams + P25CU_WALK_WARPS - 1) / P25CU_WALK_WARPS;
    if (max_blocks && blocks > max_blocks) blocks = max_blocks;   // persistent: one small CTA per SM, streams in several passes
    if (prefilter)
        p25_walk_kernel<true><<<blocks, 32 * P25CU_WALK_WARPS, 0, st>>>(p);
    else
        p25_walk_kernel<false><<<blocks, 32 * P25CU_WALK_WARPS, 0, st>>>(p);
    return cudaGetLastError();
}

// ---------------------------------------------------------------- event compaction
// Per-stream event and word counts -> exclusive offsets (one CTA; S <= a few hundred thousand).  With a word capacity
// (packed drain into a bounded host buffer) the totals cover only the leading streams that fit; the rest stay queued.
__global__ void __launch_bounds__(1024) p25_event_scan_kernel(const WalkStateHbm* states, unsigned n_streams, unsigned* offsets,
                                                              unsigned long long cap_words, unsigned* totals_out) {
    __shared__ unsigned warp_tot[2][32];
    __shared__ unsigned carry[2];
    __shared__ unsigned fit_ev, fit_w, trunc;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry[0] = carry[1] = fit_ev = fit_w = trunc = 0;
    int ovf = 0;
    __syncthreads();
    for (unsigned base = 0; base < n_streams; base += 1024) {
        const unsigned s = base + threadIdx.x;
        unsigned v[2] = {0, 0};
        if (s < n_streams) {
            v[0] = states[s].h.n_events;
            v[1] = states[s].h.n_words;
            ovf |= (int)states[s].h.overflow;
        }
        unsigned x[2] = {v[0], v[1]};
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y0 = __shfl_up_sync(FULL, x[0], o), y1 = __shfl_up_sync(FULL, x[1], o);
            if (lane >= o) {
                x[0] += y0;
                x[1] += y1;
            }
        }
        if (lane == 31) {
            warp_tot[0][warp] = x[0];
            warp_tot[1][warp] = x[1];
        }
        __syncthreads();
        if (warp < 2) {
            unsigned t = warp_tot[warp][lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned y = __shfl_up_sync(FULL, t, o);
                if (lane >= o) t += y;
            }
            warp_tot[warp][lane] = t;  // inclusive totals
        }
        __syncthreads();
        const unsigned before_ev = carry[0] + (warp ? warp_tot[0][warp - 1] : 0) + (x[0] - v[0]);
        const unsigned before_w = carry[1] + (warp ? warp_tot[1][warp - 1] : 0) + (x[1] - v[1]);
        if (s < n_streams) {
            offsets[s] = before_ev;
            offsets[n_streams + s] = before_w;
            if ((unsigned long long)before_w + v[1] <= cap_words) {
                if (v[0]) {
                    atomicMax(&fit_ev, before_ev + v[0]);
                    atomicMax(&fit_w, before_w + v[1]);
                }
            } else {
                trunc = 1;
            }
        }
        __syncthreads();
        if (threadIdx.x == 1023) {
            carry[0] = before_ev + v[0];
            carry[1] = before_w + v[1];
        }
        __syncthreads();
    }
    ovf = __syncthreads_or(ovf);
    if (threadIdx.x == 0) {
        unsigned* t = offsets + 2 * (size_t)n_streams;
        t[0] = trunc ? fit_ev : carry[0];
        t[1] = trunc ? fit_w : carry[1];
        t[2] = ovf ? 1u : 0u;
        t[3] = trunc;
        if (totals_out) {
            totals_out[0] = t[0];
            totals_out[1] = t[1];
            totals_out[2] = t[2];
            totals_out[3] = t[3];
        }
    }
}

// One warp per stream expands its packed records into 80-byte p25cu_event records and clears the slot counters.
__global__ void __launch_bounds__(256) p25_event_expand_kernel(WalkStateHbm* states, const unsigned* slots, unsigned ev_cap,
                                                               unsigned n_streams, const unsigned* offsets, p25cu_event* dense) {
    const unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (s >= n_streams) return;
    const unsigned n = states[s].h.n_events;
    const unsigned* src = slots + (size_t)s * ev_cap * P25CU_SLOT_WORDS;
    unsigned* dst = (unsigned*)(dense + offsets[s]);
    unsigned cur = 0;
    for (unsigned i = 0; i < n; i++) {
        const unsigned w1 = src[cur + 1], w2 = src[cur + 2];      // uniform (broadcast) loads
        const unsigned len = w2 >> 24, nw = p25cu_packed_words(len);
        unsigned w = 0;
        if (lane == 0) w = s;
        else if (lane == 1) w = (w2 >> 16) & 0xFFu;               // kind
        else if (lane == 2) w = w1;                               // sample, low half
        else if (lane == 3) w = w2 & 0xFFFFu;                     // sample, bits 32..47
        else if (lane == 4) w = len;
        else if (lane < 20 && (unsigned)(lane - 5 + 3) < nw) w = src[cur + lane - 2];
        if (lane < 20) dst[20 * i + lane] = w;
        cur += nw;
    }
    __syncwarp();
    if (lane == 0) {
        states[s].h.n_events = 0;
        states[s].h.n_words = 0;
        states[s].h.overflow = 0;
    }
}

// One warp per stream copies its packed words to the (stream, sample)-ordered output -- device memory or mapped
// pinned host memory, in which case the stores themselves are the device-to-host transfer -- and clears the counters.
// Streams beyond the output's capacity stay queued for the next drain.
__global__ void __launch_bounds__(256) p25_event_pack_kernel(WalkStateHbm* states, const unsigned* slots, unsigned ev_cap,
                                                             unsigned n_streams, const unsigned* offsets, unsigned* out,
                                                             unsigned long long cap_words) {
    const unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (s >= n_streams) return;
    const unsigned n = states[s].h.n_words, off = offsets[n_streams + s];
    if ((unsigned long long)off + n > cap_words) return;
    const unsigned* src = slots + (size_t)s * ev_cap * P25CU_SLOT_WORDS;
    for (unsigned i = lane; i < n; i += 32) out[off + i] = src[i];
    __syncwarp();
    if (lane == 0) {
        states[s].h.n_events = 0;
        states[s].h.n_words = 0;
        states[s].h.overflow = 0;
    }
}

cudaError_t p25cu_launch_compact(WalkStateHbm* states, const unsigned* slots, unsigned ev_cap, unsigned n_streams,
                                 unsigned* offsets, int mode, p25cu_event* dense80, unsigned* packed, unsigned long long cap_words,
                                 unsigned* totals_out, cudaStream_t st) {
    p25_event_scan_kernel<<<1, 1024, 0, st>>>(states, n_streams, offsets, mode == 2 ? cap_words : ~0ull, totals_out);
    if (mode == 0) return cudaGetLastError();  // count only (p25cu_pending)
    const unsigned blocks = (unsigned)(((size_t)n_streams * 32 + 255) / 256);
    if (mode == 1) p25_event_expand_kernel<<<blocks, 256, 0, st>>>(states, slots, ev_cap, n_streams, offsets, dense80);
    else p25_event_pack_kernel<<<blocks, 256, 0, st>>>(states, slots, ev_cap, n_streams, offsets, packed, cap_words);
    return cudaGetLastError();
}

// ---------------------------------------------------------------- FEC unit kernels (parity tests)
// One kernel per decoder (template on KIND) so that each has its own stack frame.
template <int KIND>
__global__ void p25_fec_selftest_kernel(const P25DevTables* tables, void* words, size_t count, int n, int k, void* out_data,
                                        int32_t* out_nerr) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const P25DevTables& T = *tables;
    if constexpr (KIND <= 6) {
        unsigned d = 0;
        int r;
        if constexpr (KIND == 0) r = p25_bch_decode(T, ((const unsigned long long*)words)[i], &d);
        else if constexpr (KIND == 1) r = p25_golay23_decode(T, ((const unsigned*)words)[i], &d);
        else if constexpr (KIND == 2) r = p25_golay24_decode(T, ((const unsigned*)words)[i], &d);
        else if constexpr (KIND == 3) r = p25_golay18_decode(T, ((const unsigned*)words)[i], &d);
        else if constexpr (KIND == 4) r = p25_hamming15_decode(T, ((const unsigned*)words)[i], &d);
        else if constexpr (KIND == 5) r = p25_hamming10_decode(T, ((const unsigned*)words)[i], &d);
        else r = p25_cyclic16_decode(T, ((const unsigned*)words)[i], &d);
        out_nerr[i] = r;
        ((unsigned*)out_data)[i] = d;
    } else if constexpr (KIND == 7) {
        out_nerr[i] = p25_rs_decode(T, (unsigned char*)words + i * n, n, k);  // in place, in global memory
    } else if constexpr (KIND == 8) {
        unsigned char out[12];
        out_nerr[i] = p25_trellis_half_decode(T, (const unsigned char*)words + i * 98, out);
        for (int j = 0; j < 12; j++) ((unsigned char*)out_data)[i * 12 + j] = out_nerr[i] < 0 ? 0 : out[j];
    } else if constexpr (KIND == 14) {
        unsigned char out[18];
        out_nerr[i] = p25_trellis_34_decode(T, (const unsigned char*)words + i * 98, out);
        for (int j = 0; j < 18; j++) ((unsigned char*)out_data)[i * 18 + j] = out_nerr[i] < 0 ? 0 : out[j];
    } else {
        unsigned pl[15];
        p25_imbe_decode(T, (const unsigned char*)words + i * 72, pl, pl + 8);
        for (int j = 0; j < 15; j++) ((unsigned*)out_data)[i * 15 + j] = pl[j];
        out_nerr[i] = 0;
    }
}

// The warp-cooperative decoders the walker uses, one warp per word: kinds 10 RS, 11 IMBE, 12 BCH, 13 half-rate trellis,
// 15 3/4-rate trellis (same inputs and outputs as kinds 7, 9, 0, 8, 14).
__global__ void __launch_bounds__(32 * P25CU_WALK_WARPS) p25_fec_selftest_warp_kernel(const P25DevTables* tables, int kind, void* words,
                                                                                      size_t count, int n, int k, void* out_data,
                                                                                      int32_t* out_nerr) {
    __shared__ WalkShared sh;
    walk_shared_init(sh, tables);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t i = (size_t)blockIdx.x * P25CU_WALK_WARPS + warp;
    if (i >= count) return;
    unsigned char* scr = sh.scr[warp];
    unsigned char* buf = sh.ws[warp].buf;
    if (kind == 10) {
        unsigned char* sym = (unsigned char*)words + i * n;
        if (lane < n) buf[lane] = sym[lane];
        if (lane + 32 < n) buf[lane + 32] = sym[lane + 32];
        __syncwarp();
        const int r = warp_rs_decode(sh.T, scr, buf, n, k, lane);
        __syncwarp();
        if (lane < n) sym[lane] = buf[lane];
        if (lane + 32 < n) sym[lane + 32] = buf[lane + 32];
        if (lane == 0) out_nerr[i] = r;
    } else if (kind == 11) {
        for (int j = lane; j < 72; j += 32) buf[j] = ((const unsigned char*)words)[i * 72 + j];
        __syncwarp();
        unsigned* pl = reinterpret_cast<unsigned*>(scr + 128);
        warp_imbe_decode(sh, buf, lane, pl);
        if (lane < 15) ((unsigned*)out_data)[i * 15 + lane] = pl[lane];
        if (lane == 0) out_nerr[i] = 0;
    } else if (kind == 12) {
        unsigned d = 0;
        const int r = warp_bch_decode(sh.T, scr, ((const unsigned long long*)words)[i], lane, &d);
        if (lane == 0) {
            out_nerr[i] = r;
            ((unsigned*)out_data)[i] = d;
        }
    } else if (kind == 15) {
        for (int j = lane; j < 98; j += 32) buf[j] = ((const unsigned char*)words)[i * 98 + j];
        __syncwarp();
        unsigned char* out = scr + 128;
        const int r = warp_trellis_34_decode(sh.T, sh.surv[warp], buf, lane, out);
        __syncwarp();
        if (lane < 18) ((unsigned char*)out_data)[i * 18 + lane] = r < 0 ? 0 : out[lane];
        if (lane == 0) out_nerr[i] = r;
    } else {
        for (int j = lane; j < 98; j += 32) buf[j] = ((const unsigned char*)words)[i * 98 + j];
        __syncwarp();
        unsigned char* out = sh.ws[warp].hex;
        const int r = warp_trellis_half_decode(sh.T, scr, buf, lane, out);
        __syncwarp();
        if (lane < 12) ((unsigned char*)out_data)[i * 12 + lane] = r < 0 ? 0 : out[lane];
        if (lane == 0) out_nerr[i] = r;
    }
}

cudaError_t p25cu_launch_fec_selftest(const P25DevTables* tables, int kind, void* words, size_t count, int n, int k,
                                      void* out_data, int32_t* out_nerr, cudaStream_t st) {
    if (count == 0) return cudaSuccess;
    if ((kind >= 10 && kind <= 13) || kind == 15) {
        const unsigned wb = (unsigned)((count + P25CU_WALK_WARPS - 1) / P25CU_WALK_WARPS);
        p25_fec_selftest_warp_kernel<<<wb, 32 * P25CU_WALK_WARPS, 0, st>>>(tables, kind, words, count, n, k, out_data, out_nerr);
        return cudaGetLastError();
    }
    const unsigned blocks = (unsigned)((count + 127) / 128);
    if (blocks == 0) return cudaSuccess;
#define P25_ST_CASE(K) case K: p25_fec_selftest_kernel<K><<<blocks, 128, 0, st>>>(tables, words, count, n, k, out_data, out_nerr); break;
    switch (kind) {
        P25_ST_CASE(0) P25_ST_CASE(1) P25_ST_CASE(2) P25_ST_CASE(3) P25_ST_CASE(4)
        P25_ST_CASE(5) P25_ST_CASE(6) P25_ST_CASE(7) P25_ST_CASE(8) P25_ST_CASE(9) P25_ST_CASE(14)
        default: return cudaErrorInvalidValue;
    }
#undef P25_ST_CASE
    return cudaGetLastError();
}
