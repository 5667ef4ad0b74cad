// decode_walk.cu -- frame sync, symbol slicing, deframing and FEC decode for many streams.
//
// Replaces, per stream, the loop `for &s in samples { msg.feed(s) }` of the reference
// (src/recv.rs:148-150, :204-234; src/replay.rs:40-57) where `msg` is p25::MessageReceiver
// (crate p25 @a96c564, not vendored).  One warp owns one stream.  Instead of stepping one
// sample at a time the warp advances in bulk between "decision points":
//
//   SYNC    32 candidate positions per step; every lane evaluates the 231-tap frame-sync
//           correlation and window energy of one position, neighbouring results travel by
//           warp shuffle and the first lane whose predicate fires is found with a ballot.
//   symbols 32 symbol instants per step (stride 10 samples); lanes slice their sample
//           against the three thresholds, a ballot removes status symbols (every 36th
//           dibit) and gives each data dibit its index in the unit buffer.
//   unit    when the dibit that completes a NID / TSBK block / voice frame / ... has been
//           stored, lane 0 runs the FEC decoder and writes the event; state is shared
//           through the warp's shared-memory record.
//
// Every step is an exact restatement of the sample-at-a-time rule (oracle/p25_oracle.cpp),
// so events and their sample indices are bit-identical for any chunking.  This file is
// compiled with --fmad=false: the only fused operations are the explicit fmaf() of the
// correlator, in the same order as the oracle.
#include "p25cu_internal.cuh"

#define FULL 0xFFFFFFFFu

__constant__ float c_sync_fp[P25_FP_LEN];

struct WalkShared {
    P25DevTables T;
    WalkState ws[P25CU_WALK_WARPS];
};
static_assert(sizeof(P25DevTables) % 16 == 0, "tables are staged with 16-byte copies");

struct WarpCtx {
    const WalkParams* p;
    const P25DevTables* T;
    WalkState* ws;
    unsigned stream;
};

// ---------------------------------------------------------------- lane-0 helpers
__device__ __forceinline__ void stat_ok(const WarpCtx& c, int fam, unsigned fixed) {
    unsigned* s = c.p->stats + ((size_t)c.stream * P25CU_ST_FAMILIES + fam) * 3;
    s[0] += 1;
    s[2] += fixed;
}
__device__ __forceinline__ void stat_bad(const WarpCtx& c, int fam) {
    unsigned* s = c.p->stats + ((size_t)c.stream * P25CU_ST_FAMILIES + fam) * 3;
    s[0] += 1;
    s[1] += 1;
}

__device__ __noinline__ void emit(const WarpCtx& c, unsigned kind, unsigned long long idx, const void* payload, unsigned len) {
    WalkState& ws = *c.ws;
    if (ws.n_events >= c.p->ev_cap) {
        ws.overflow = 1;
        return;
    }
    union {
        p25cu_event e;
        uint4 q[5];
    } u;
    for (int i = 0; i < 5; i++) u.q[i] = make_uint4(0, 0, 0, 0);
    u.e.stream = c.stream;
    u.e.kind = kind;
    u.e.sample = idx;
    u.e.len = len;
    const unsigned char* src = (const unsigned char*)payload;
    for (unsigned i = 0; i < len; i++) u.e.payload[i] = src[i];
    uint4* dst = (uint4*)(c.p->slots + (size_t)c.stream * c.p->ev_cap + ws.n_events);
    for (int i = 0; i < 5; i++) dst[i] = u.q[i];
    ws.n_events++;
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

// NID complete (32 dibits in ws.buf).  PacketNID event, reference src/recv.rs:216-222.
__device__ __noinline__ void complete_nid(const WarpCtx& c, unsigned long long idx) {
    WalkState& ws = *c.ws;
    const P25DevTables& T = *c.T;
    unsigned long long bits = 0;
    for (int i = 0; i < P25_NID_DIBITS; i++) bits = (bits << 2) | ws.buf[i];
    unsigned data;
    const int nerr = p25_bch_decode(T, bits >> 1, &data);
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
            ws.state = WS_PAYLOAD;
            break;
        case 0x3:
            ws.state = WS_FLUSH;
            break;
        case 0xC:
            enter_sync(ws, idx + 1);
            break;
        default:
            fail(c, P25CU_E_UNKNOWN_NID, idx);
            return;
    }
    const unsigned char pl[3] = {(unsigned char)((data >> 4) & 0xFF), (unsigned char)(data >> 12), (unsigned char)ws.duid};
    emit(c, P25CU_EV_NID, idx, pl, 3);
}

// A payload decision point has been reached (ws.cnt == target).
__device__ __noinline__ void complete_payload(const WarpCtx& c, unsigned long long idx) {
    WalkState& ws = *c.ws;
    const P25DevTables& T = *c.T;
    switch (ws.duid) {
        case 0x7: {  // TSBK -> TrunkingControl (reference src/recv.rs:231)
            unsigned char out[12];
            const int fixed = p25_trellis_half_decode(T, ws.buf, out);
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
            return;
        }
        case 0x0: {  // HDU -> VoiceHeader (reference src/recv.rs:223)
            for (int w = 0; w < 36; w++) {
                unsigned d6;
                const int n = p25_golay18_decode(T, p25_take_bits(ws.buf, 18 * w, 18), &d6);
                if (n < 0) stat_bad(c, P25CU_ST_GOLAY_SHORT); else stat_ok(c, P25CU_ST_GOLAY_SHORT, (unsigned)n);
                ws.hex[w] = (unsigned char)d6;
            }
            const int n = p25_rs_decode(T, ws.hex, 36, 20);
            if (n < 0) {
                stat_bad(c, P25CU_ST_RS_LONG);
                fail(c, P25CU_E_RS, idx);
                return;
            }
            stat_ok(c, P25CU_ST_RS_LONG, (unsigned)n);
            unsigned char out[15];
            p25_pack_hexbits(ws.hex, 20, out);
            ws.state = WS_FLUSH;
            emit(c, P25CU_EV_VOICE_HEADER, idx, out, 15);
            return;
        }
        case 0xF: {  // TDULC -> VoiceTerm (reference src/recv.rs:232)
            for (int w = 0; w < 12; w++) {
                unsigned d12;
                const unsigned word = p25_take_bits(ws.buf, 24 * w, 24);
                const int n = p25_golay24_decode(T, word, &d12);
                if (n < 0) {
                    stat_bad(c, P25CU_ST_GOLAY_EXT);
                    d12 = word >> 12;
                } else {
                    stat_ok(c, P25CU_ST_GOLAY_EXT, (unsigned)n);
                }
                ws.hex[2 * w] = (unsigned char)(d12 >> 6);
                ws.hex[2 * w + 1] = (unsigned char)(d12 & 0x3F);
            }
            const int n = p25_rs_decode(T, ws.hex, 24, 12);
            if (n < 0) {
                stat_bad(c, P25CU_ST_RS_SHORT);
                fail(c, P25CU_E_RS, idx);
                return;
            }
            stat_ok(c, P25CU_ST_RS_SHORT, (unsigned)n);
            unsigned char out[9];
            p25_pack_hexbits(ws.hex, 12, out);
            ws.state = WS_FLUSH;
            emit(c, P25CU_EV_VOICE_TERM, idx, out, 9);
            return;
        }
        default: {  // LDU1 / LDU2 parts (reference src/recv.rs:224-230)
            const int kind = T.ldu_kind[ws.part];
            const unsigned char* d = ws.buf + T.ldu_start[ws.part];
            ws.part++;
            if (kind == 0) {
                unsigned pl[15];
                p25_imbe_decode(T, d, pl, pl + 8);
                for (int i = 0; i < 4; i++) stat_ok(c, P25CU_ST_GOLAY_STD, pl[8 + i]);
                for (int i = 4; i < 7; i++) stat_ok(c, P25CU_ST_HAMMING_STD, pl[8 + i]);
                if (ws.part == P25_LDU_PARTS) ws.state = WS_FLUSH;
                emit(c, P25CU_EV_VOICE_FRAME, idx, pl, 60);
                return;
            }
            if (kind == 1) {
                for (int w = 0; w < 4; w++) {
                    unsigned d6;
                    const int n = p25_hamming10_decode(T, p25_take_bits(d, 10 * w, 10), &d6);
                    if (n < 0) stat_bad(c, P25CU_ST_HAMMING_SHORT); else stat_ok(c, P25CU_ST_HAMMING_SHORT, (unsigned)n);
                    ws.hex[4 * ws.chunks + w] = (unsigned char)d6;
                }
                if (++ws.chunks < 6) return;
                const bool lc = ws.duid == 0x5;
                const int n = p25_rs_decode(T, ws.hex, 24, lc ? 12 : 16);
                if (n < 0) {
                    stat_bad(c, lc ? P25CU_ST_RS_SHORT : P25CU_ST_RS_MED);
                    fail(c, P25CU_E_RS, idx);
                    return;
                }
                stat_ok(c, lc ? P25CU_ST_RS_SHORT : P25CU_ST_RS_MED, (unsigned)n);
                unsigned char out[12];
                p25_pack_hexbits(ws.hex, lc ? 12 : 16, out);
                emit(c, lc ? P25CU_EV_LINK_CONTROL : P25CU_EV_CRYPTO_CONTROL, idx, out, lc ? 9 : 12);
                return;
            }
            unsigned frag = 0;
            for (int w = 0; w < 2; w++) {
                unsigned d8;
                const int n = p25_cyclic16_decode(T, p25_take_bits(d, 16 * w, 16), &d8);
                if (n < 0) stat_bad(c, P25CU_ST_CYCLIC); else stat_ok(c, P25CU_ST_CYCLIC, (unsigned)n);
                frag = (frag << 8) | d8;
            }
            emit(c, P25CU_EV_LSD, idx, &frag, 4);
            return;
        }
    }
}

// ---------------------------------------------------------------- the walker
__global__ void __launch_bounds__(32 * P25CU_WALK_WARPS) p25_walk_kernel(const WalkParams p) {
    __shared__ WalkShared sh;
    {
        const uint4* src = (const uint4*)p.tables;
        uint4* dst = (uint4*)&sh.T;
        for (unsigned i = threadIdx.x; i < sizeof(P25DevTables) / 16; i += blockDim.x) dst[i] = src[i];
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned stream = blockIdx.x * P25CU_WALK_WARPS + warp;
    const bool active = stream < p.n_streams;
    WalkState& ws = sh.ws[warp];
    if (active) {
        const uint4* src = (const uint4*)(p.states + stream);
        uint4* dst = (uint4*)&ws;
        for (unsigned i = lane; i < sizeof(WalkState) / 16; i += 32) dst[i] = src[i];
    }
    __syncthreads();
    if (!active) return;

    WarpCtx c{&p, &sh.T, &ws, stream};
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
            // ---- frame-sync search over 32 positions
            const unsigned long long pos = ws.pos;
            if (pos >= end) break;
            const unsigned long long n_abs = pos + lane;
            const bool valid = n_abs < end;
            float corr = 0.f, en = 0.f;
            if (valid) {
                const float* w = row + (long long)(n_abs - p0) + P25CU_BB_HIST - (P25_FP_LEN - 1);
#pragma unroll
                for (int k = 0; k < P25_FP_LEN; k++) {
                    const float x = __ldg(w + k);
                    corr = fmaf(c_sync_fp[k], x, corr);
                    en = fmaf(x, x, en);
                }
            }
            const bool above = valid && corr > 0.f && (corr * corr >= P25_SYNC_RHO2_EFP * en);
            float pc = __shfl_up_sync(FULL, corr, 1);
            int pa = __shfl_up_sync(FULL, (int)above, 1);
            int hp = 1;
            if (lane == 0) {
                pc = ws.prev_corr;
                pa = ws.prev_above;
                hp = ws.have_prev;
            }
            const bool fire = valid && hp && above && pa && corr <= pc;
            const unsigned fm = __ballot_sync(FULL, fire);
            if (fm == 0) {
                const unsigned long long left = end - pos;
                const int nvalid = left < 32 ? (int)left : 32;
                const float lc = __shfl_sync(FULL, corr, nvalid - 1);
                const int la = __shfl_sync(FULL, (int)above, nvalid - 1);
                __syncwarp();
                if (lane == 0) {
                    ws.prev_corr = lc;
                    ws.prev_above = la;
                    ws.have_prev = 1;
                    ws.pos = pos + nvalid;
                }
                __syncwarp();
                continue;
            }
            // ---- lock: the correlation peaked on the sample before the firing one
            const unsigned long long idx = pos + (__ffs(fm) - 1);
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
            if (duid == 0x7) target = P25_TSBK_DIBITS;
            else if (duid == 0x0) target = P25_HDU_DIBITS;
            else if (duid == 0xF) target = P25_TDULC_DIBITS;
            else target = sh.T.ldu_start[ws.part] + sh.T.ldu_len[ws.part];
        } else {
            target = 0x7FFFFFFF;
        }
        const int need = target - cnt;
        const unsigned long long span = (end - 1 - t0) / P25_SPS + 1;
        const int avail = span < 32 ? (int)span : 32;
        const bool valid = lane < avail;
        const bool status = ((fpos0 + lane) % P25_STATUS_PERIOD) == P25_STATUS_PERIOD - 1;
        float s = 0.f;
        if (valid) s = __ldg(row + (long long)(t0 - p0) + P25CU_BB_HIST + P25_SPS * lane);
        const int d = s > ws.pth ? 1 : (s > ws.mid ? 0 : (s > ws.nth ? 2 : 3));  // [STD] 01 +3, 00 +1, 10 -1, 11 -3
        const unsigned dmask = __ballot_sync(FULL, valid && !status);
        const int c_incl = __popc(dmask & lane_le);
        const unsigned stopmask = (st == WS_FLUSH) ? __ballot_sync(FULL, valid && status)
                                                   : __ballot_sync(FULL, valid && !status && c_incl == need);
        const int K = stopmask ? __ffs(stopmask) : avail;
        if (st != WS_FLUSH && lane < K && valid && !status) ws.buf[cnt + c_incl - 1] = (unsigned char)d;
        const int ndata = __popc(dmask & (K >= 32 ? FULL : ((1u << K) - 1u)));
        __syncwarp();
        if (lane == 0) {
            ws.next_sym = t0 + (unsigned long long)P25_SPS * K;
            ws.frame_pos = fpos0 + K;
            if (st != WS_FLUSH) ws.cnt = cnt + ndata;
            if (stopmask) {
                const unsigned long long idx = t0 + (unsigned long long)P25_SPS * (K - 1);
                if (st == WS_FLUSH) enter_sync(ws, idx + 1);
                else if (st == WS_NID) complete_nid(c, idx);
                else complete_payload(c, idx);
            }
        }
        __syncwarp();
    }

    // ---- end of chunk: persist state, roll the last 256 samples to the front of the row
    __syncwarp();
    {
        uint4* dst = (uint4*)(p.states + stream);
        const uint4* src = (const uint4*)&ws;
        for (unsigned i = lane; i < sizeof(WalkState) / 16; i += 32) dst[i] = src[i];
    }
    float* rw = p.bb_rw + (size_t)stream * p.row_stride;
    float keep[P25CU_BB_HIST / 32];
#pragma unroll
    for (int i = 0; i < P25CU_BB_HIST / 32; i++) keep[i] = rw[p.n + lane + 32 * i];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < P25CU_BB_HIST / 32; i++) rw[lane + 32 * i] = keep[i];
}

cudaError_t p25cu_walk_upload_consts() {
    return cudaMemcpyToSymbol(c_sync_fp, P25_SYNC_FP, sizeof(float) * P25_FP_LEN);
}

cudaError_t p25cu_launch_walk(const WalkParams& p, cudaStream_t st) {
    const unsigned blocks = (p.n_streams + P25CU_WALK_WARPS - 1) / P25CU_WALK_WARPS;
    p25_walk_kernel<<<blocks, 32 * P25CU_WALK_WARPS, 0, st>>>(p);
    return cudaGetLastError();
}

// ---------------------------------------------------------------- event compaction
// Per-stream slot counts -> exclusive offsets (one CTA; S <= a few hundred thousand).
__global__ void __launch_bounds__(1024) p25_event_scan_kernel(const WalkState* states, unsigned n_streams, unsigned* offsets) {
    __shared__ unsigned warp_tot[32];
    __shared__ unsigned carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    int ovf = 0;
    __syncthreads();
    for (unsigned base = 0; base < n_streams; base += 1024) {
        const unsigned s = base + threadIdx.x;
        const unsigned v = s < n_streams ? states[s].n_events : 0;
        if (s < n_streams) ovf |= (int)states[s].overflow;
        unsigned x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(FULL, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_tot[warp] = x;
        __syncthreads();
        if (warp == 0) {
            unsigned t = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned y = __shfl_up_sync(FULL, t, o);
                if (lane >= o) t += y;
            }
            warp_tot[lane] = t;  // inclusive totals
        }
        __syncthreads();
        const unsigned before = carry + (warp ? warp_tot[warp - 1] : 0) + (x - v);
        if (s < n_streams) offsets[s] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    ovf = __syncthreads_or(ovf);
    if (threadIdx.x == 0) {
        offsets[n_streams] = carry;
        offsets[n_streams + 1] = ovf ? 1u : 0u;
    }
}

// One warp per stream copies its events to the dense array and clears the slot count.
__global__ void __launch_bounds__(256) p25_event_gather_kernel(WalkState* states, const p25cu_event* slots, unsigned ev_cap,
                                                               unsigned n_streams, const unsigned* offsets, p25cu_event* dense) {
    const unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (s >= n_streams) return;
    const unsigned n = states[s].n_events;
    const uint4* src = (const uint4*)(slots + (size_t)s * ev_cap);
    uint4* dst = (uint4*)(dense + offsets[s]);
    for (unsigned i = lane; i < n * 5; i += 32) dst[i] = src[i];
    __syncwarp();
    if (lane == 0) {
        states[s].n_events = 0;
        states[s].overflow = 0;
    }
}

cudaError_t p25cu_launch_compact(const WalkState* states, const p25cu_event* slots, unsigned ev_cap, unsigned n_streams,
                                 unsigned* offsets, p25cu_event* dense, cudaStream_t st) {
    p25_event_scan_kernel<<<1, 1024, 0, st>>>(states, n_streams, offsets);
    if (!dense) return cudaGetLastError();  // count only (p25cu_pending)
    const unsigned blocks = (n_streams * 32 + 255) / 256;
    p25_event_gather_kernel<<<blocks, 256, 0, st>>>(const_cast<WalkState*>(states), slots, ev_cap, n_streams, offsets, dense);
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
    } else {
        unsigned pl[15];
        p25_imbe_decode(T, (const unsigned char*)words + i * 72, pl, pl + 8);
        for (int j = 0; j < 15; j++) ((unsigned*)out_data)[i * 15 + j] = pl[j];
        out_nerr[i] = 0;
    }
}

cudaError_t p25cu_launch_fec_selftest(const P25DevTables* tables, int kind, void* words, size_t count, int n, int k,
                                      void* out_data, int32_t* out_nerr, cudaStream_t st) {
    const unsigned blocks = (unsigned)((count + 127) / 128);
    if (blocks == 0) return cudaSuccess;
#define P25_ST_CASE(K) case K: p25_fec_selftest_kernel<K><<<blocks, 128, 0, st>>>(tables, words, count, n, k, out_data, out_nerr); break;
    switch (kind) {
        P25_ST_CASE(0) P25_ST_CASE(1) P25_ST_CASE(2) P25_ST_CASE(3) P25_ST_CASE(4)
        P25_ST_CASE(5) P25_ST_CASE(6) P25_ST_CASE(7) P25_ST_CASE(8) P25_ST_CASE(9)
        default: return cudaErrorInvalidValue;
    }
#undef P25_ST_CASE
    return cudaGetLastError();
}
