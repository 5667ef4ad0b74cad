// pfb.cu -- K1b: wideband polyphase channelizer + per-channel C4FM baseband (BASELINE.json configs[2]).
//
// No stage of the reference does this (one RTL-SDR tuner, one channel, retuned by hopping: reference
// src/sdr.rs:61-68, src/recv.rs:127-137); it is the declared extension SURVEY.md section 2.5 names K1b.  A
// 19.2 MS/s cf32 capture is split into 1,536 channels 12.5 kHz apart, each at 48 kS/s, and every channel
// then runs the reference's own 48 kHz stages:
//   channel-select FIR   reference src/demod.rs:93      (static_fir::FirFilter<BandpassFir>)
//   power accumulation   reference src/demod.rs:95-101
//   FM discriminator     reference src/demod.rs:109-111
//   10-tap moving average reference src/demod.rs:114
// so that the decode walker sees 1,536 ordinary streams per capture.
//
// Channel k is defined (spec/p25_spec.py) as mix-down by k * 12.5 kHz, a 6,144-tap prototype low-pass and
// decimation by 400:   y_k[m] = sum_i h[i] x[n_m - i] exp(-2j pi k (n_m - i) / 1536),  n_m = 400 m + 399.
// Writing i = r + 1536 p gives the polyphase form this file computes:
//   v_m[r] = sum_{p<4} h[r + 1536 p] x[n_m - r - 1536 p]           (4 complex-by-real taps per branch)
//   y_k[m] = sum_q v_m[(q + n_m) mod 1536] exp(+2j pi k q / 1536)   (one 1,536-point inverse DFT per output time)
//
// Round 2 folds the channel-select FIR into the prototype (see kernel A below): kernel A = polyphase branches + FFT +
// discriminator, kernel B = boxcar + transpose.  Algorithmic HBM traffic: 8 B per input sample in, 4 B per channel and
// output time out (15.4 B per input sample); the discriminator rows between the two kernels add 2 x 15.4 B.
#include "p25cu_internal.cuh"
#include "p25_pfb_taps.h"

namespace pfb {

constexpr int N = P25_PFB_N, M = P25_PFB_M, P = P25_PFB_P, L = N * P;
constexpr int NT = 256;
static_assert(N == 3 * 8 * 8 * 8, "FFT plan is 3 x 8 x 8 x 8");

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_i(float2 a) { return make_float2(-a.y, a.x); }     // a * (+i)
// acc += h * x on both components: one FFMA2
__device__ __forceinline__ float2 cfma(float h, float2 x, float2 acc) {
    unsigned long long r;
    const float2 hh = make_float2(h, h);
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(r)
        : "l"(*reinterpret_cast<const unsigned long long*>(&hh)), "l"(*reinterpret_cast<const unsigned long long*>(&x)),
          "l"(*reinterpret_cast<const unsigned long long*>(&acc)));
    return *reinterpret_cast<float2*>(&r);
}

// 8-point DFT with the + sign: w[q] = sum_r v[r] exp(+2 pi i r q / 8), in place
__device__ __forceinline__ void dft8(float2 (&v)[8]) {
    const float h = 0.70710678118654752f;
    // radix-2 decimation in time: evens (0,2,4,6), odds (1,3,5,7)
    float2 e0 = cadd(v[0], v[4]), e1 = csub(v[0], v[4]), e2 = cadd(v[2], v[6]), e3 = mul_i(csub(v[2], v[6]));
    float2 o0 = cadd(v[1], v[5]), o1 = csub(v[1], v[5]), o2 = cadd(v[3], v[7]), o3 = mul_i(csub(v[3], v[7]));
    const float2 E0 = cadd(e0, e2), E2 = csub(e0, e2), E1 = cadd(e1, e3), E3 = csub(e1, e3);   // 4-point DFT of the evens
    const float2 O0 = cadd(o0, o2), O2 = csub(o0, o2), O1 = cadd(o1, o3), O3 = csub(o1, o3);   // ... of the odds
    // odd outputs times exp(+2 pi i q / 8), q = 0..3
    const float2 T0 = O0;
    const float2 T1 = make_float2(h * (O1.x - O1.y), h * (O1.x + O1.y));     // * (1 + i) / sqrt 2
    const float2 T2 = mul_i(O2);
    const float2 T3 = make_float2(-h * (O3.x + O3.y), h * (O3.x - O3.y));    // * (-1 + i) / sqrt 2
    v[0] = cadd(E0, T0);
    v[4] = csub(E0, T0);
    v[1] = cadd(E1, T1);
    v[5] = csub(E1, T1);
    v[2] = cadd(E2, T2);
    v[6] = csub(E2, T2);
    v[3] = cadd(E3, T3);
    v[7] = csub(E3, T3);
}

__device__ __forceinline__ float atan2_branchfree(float y, float x) {   // same polynomial as ddc_fm.cu disc_atan2
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    float rc;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(fmaxf(mx, 1e-30f)));
    const float t = mn * rc, u = t * t;
    float q = -0.004295386839658022f;
    q = fmaf(q, u, 0.022737378254532814f);
    q = fmaf(q, u, -0.057179518043994904f);
    q = fmaf(q, u, 0.09735459089279175f);
    q = fmaf(q, u, -0.13945257663726807f);
    q = fmaf(q, u, 0.1995391547679901f);
    q = fmaf(q, u, -0.3333050608634949f);
    q = fmaf(q, u, 0.9999995231628418f);
    float r = q * t;
    r = ay > ax ? 1.57079632679489662f - r : r;
    r = x < 0.f ? 3.14159265358979324f - r : r;
    return copysignf(r, y);
}

// ------------------------------------------------------------------------------------------------ kernel A
// Round 2: the per-channel channel-select FIR (41 taps at 48 kS/s, src/demod.rs:93) is folded into the prototype.
// Channel k after the channel filter is
//     c_k[m] = sum_j hc[j] y_k[m - j] = sum_i heq[i] x[n_m - i] exp(-2j pi k (n_m - i) / N),
//     heq = hp (*) upsample(hc, M)           (6,144 + 40 * 400 = 22,144 taps, zero-padded to PE * N = 23,040)
// i.e. the same polyphase form with PE = 15 taps per branch instead of 4 -- 23,040 complex-by-real MACs per output time
// for all 1,536 channels, instead of 6,144 + 1,536 * 41 = 69,120 -- and no [time][channel] spectra round trip through HBM
// before the FIR (the round-1 design moved 61 bytes per input sample there).
//
// A CTA (512 threads, one per SM: 222 KB of shared memory) owns a run of consecutive output times of one capture:
//   * the input window lives in a shared-memory ring of 16 x 1,536 samples (192 KB) indexed by the logical sample
//     index (carried tail ++ chunk); every output time brings M = 400 new samples, prefetched one step ahead;
//   * a thread owns three polyphase branches r = tid + 512 b and keeps their 45 prototype taps in registers for the
//     whole run, so the branch sums cost one LDS.64 + one FFMA2 per tap;
//   * the 1,536-point inverse DFT is the mixed-radix Stockham FFT 3 x 8 x 8 x 8 in shared memory; its last pass leaves
//     every thread with the same eight channels at every output time, so the previous c_k[m - 1] stays in registers and
//     the FM discriminator (src/demod.rs:109-111) runs right there: only d_k[m] (4 bytes per channel and time) goes to
//     HBM, time-major; power (src/demod.rs:95-101) accumulates in registers;
//   * one warm-up output time per run supplies c_k[m0 - 1].
// Kernel B then only applies the 10-tap boxcar (src/demod.rs:114) along time and transposes into the baseband rows.
constexpr int PE = 15;                 // taps per branch of the equivalent prototype
constexpr int LE = N * PE;             // 23,040
constexpr int HTX = 23552;             // carried input tail (>= LE + M - 1, a multiple of 128)
constexpr int RROWS = 16, RING = RROWS * N;
constexpr int NTX = 512;
constexpr int DHIST = P25_BOXCAR - 1;  // discriminator rows carried in front of every chunk (9)
static_assert(HTX >= LE + M - 1 && RING >= LE + 2 * M, "window + prefetch must fit");

struct SmemX {
    float2 ring[RING];
    float2 fa[N], fb[N];
    // first twiddle of every radix-8 butterfly, laid out by the lane index of its pass (conflict-free LDS.64); the
    // other six are its powers, formed in registers (a 1,536-entry table indexed r * k * stride put up to 32 lanes
    // on one bank: 26 % of the kernel's shared-memory wavefronts in the first version)
    float2 tw3[4], tw24[24], tw192[N / 8];   // exp(+2 pi i k / 24), exp(+2 pi i k / 192), exp(+2 pi i k / 1536)
};

struct PfbParams {
    const float2* iq;          // [captures][n]
    const float2* tail_in;     // [captures][HTX]
    const float* taps;         // [LE] equivalent prototype (prototype (*) upsampled channel filter)
    const float2* twiddle;     // [N] exp(+2 pi i t / N)
    float* d;                  // [captures][d_rows][N] discriminator output, time-major; this chunk's rows start at DHIST
    float2* y;                 // nullable: [captures][y_rows][N] channel-filtered spectra c_k[m] of this chunk (test hook)
    float* power_sum;          // nullable: [captures * N]
    unsigned long long a0, m0; // absolute input / output index of the chunk start
    unsigned n, n_out, n_captures;
    unsigned d_rows, y_rows;
};

__device__ __forceinline__ float2 load_logical(const PfbParams& p, const float2* tail, const float2* chunk, long long l) {
    if (l < HTX) return l >= 0 ? tail[l] : make_float2(0.f, 0.f);
    const long long i = l - HTX;
    return i < (long long)p.n ? __ldg(chunk + i) : make_float2(0.f, 0.f);
}
// w^1 .. w^7 from w: six complex multiplies, depth three
__device__ __forceinline__ void powers7(float2 w1, float2 (&w)[8]) {
    w[1] = w1;
    w[2] = cmul(w1, w1);
    w[3] = cmul(w[2], w1);
    w[4] = cmul(w[2], w[2]);
    w[5] = cmul(w[4], w1);
    w[6] = cmul(w[3], w[3]);
    w[7] = cmul(w[6], w1);
}

// one radix-8 Stockham pass; w1 = exp(+2 pi i (j % NS) / (8 NS)) is the butterfly's first twiddle
template <int NS>
__device__ __forceinline__ void pass8x(const float2* __restrict__ in, float2* __restrict__ out, float2 w1, int j) {
    float2 v[8], w[8];
#pragma unroll
    for (int r = 0; r < 8; r++) v[r] = in[j + r * (N / 8)];
    powers7(w1, w);
#pragma unroll
    for (int r = 1; r < 8; r++) v[r] = cmul(v[r], w[r]);
    dft8(v);
    const int j0 = (j / NS) * NS * 8 + (j % NS);
#pragma unroll
    for (int r = 0; r < 8; r++) out[j0 + r * NS] = v[r];
}

// branch sum of one polyphase branch whose window starts in ring row R0 (mod 16): every row offset is a constant
template <int R0>
__device__ __forceinline__ float2 branch_sum(const float2* __restrict__ col, const float (&tap)[PE]) {
    float2 x[PE];
#pragma unroll
    for (int pp = 0; pp < PE; pp++) x[pp] = col[((R0 - pp) & (RROWS - 1)) * N];        // all 15 loads in flight first
    float2 a0 = make_float2(0.f, 0.f), a1 = a0, a2 = a0;                               // three chains of five: 20 instead of 60 cycles deep
#pragma unroll
    for (int pp = 0; pp < PE; pp += 3) {
        a0 = cfma(tap[pp], x[pp], a0);
        a1 = cfma(tap[pp + 1], x[pp + 1], a1);
        a2 = cfma(tap[pp + 2], x[pp + 2], a2);
    }
    return cadd(cadd(a0, a1), a2);
}
__device__ __forceinline__ float2 branch_sum_dyn(unsigned row0, const float2* __restrict__ col, const float (&tap)[PE]) {
    switch (row0 & (RROWS - 1)) {       // warp-uniform up to the one column where the window wraps
        case 0: return branch_sum<0>(col, tap);
        case 1: return branch_sum<1>(col, tap);
        case 2: return branch_sum<2>(col, tap);
        case 3: return branch_sum<3>(col, tap);
        case 4: return branch_sum<4>(col, tap);
        case 5: return branch_sum<5>(col, tap);
        case 6: return branch_sum<6>(col, tap);
        case 7: return branch_sum<7>(col, tap);
        case 8: return branch_sum<8>(col, tap);
        case 9: return branch_sum<9>(col, tap);
        case 10: return branch_sum<10>(col, tap);
        case 11: return branch_sum<11>(col, tap);
        case 12: return branch_sum<12>(col, tap);
        case 13: return branch_sum<13>(col, tap);
        case 14: return branch_sum<14>(col, tap);
        default: return branch_sum<15>(col, tap);
    }
}

__global__ void __launch_bounds__(NTX, 1) p25_pfbx_kernel(const PfbParams p, const unsigned times_per_cta) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemX& sm = *reinterpret_cast<SmemX*>(smem_raw);
    const int tid = threadIdx.x;
    const unsigned cap = blockIdx.y;
    const int t_first = (int)(blockIdx.x * times_per_cta);       // this CTA's consecutive output times, relative to m0
    const int t_last = min(t_first + (int)times_per_cta, (int)p.n_out);
    if (t_first >= t_last) return;
    const float2* chunk = p.iq + (size_t)cap * p.n;
    const float2* tail = p.tail_in + (size_t)cap * HTX;
    if (tid < N / 8) sm.tw192[tid] = p.twiddle[tid];
    if (tid < 24) sm.tw24[tid] = p.twiddle[8 * tid];
    if (tid < 3) sm.tw3[tid] = p.twiddle[64 * tid];
    float tap[3][PE];
#pragma unroll
    for (int b = 0; b < 3; b++)
#pragma unroll
        for (int pp = 0; pp < PE; pp++) tap[b][pp] = __ldg(p.taps + tid + NTX * b + N * pp);
    // logical index (tail ++ chunk, tail = HTX samples) of the newest input of output time t: n_m - a0 + HTX
    const long long e_base = (long long)M * (long long)p.m0 + (M - 1) - (long long)p.a0 + HTX;
    {   // the window of the warm-up output time t_first - 1
        const long long e0 = e_base + (long long)M * (t_first - 1);
        for (int i = tid; i < LE; i += NTX) {
            const long long l = e0 - (LE - 1) + i;
            sm.ring[(int)((l + 4ll * RING) % RING)] = load_logical(p, tail, chunk, l);
        }
    }
    float2 nx = make_float2(0.f, 0.f);                            // next output time's new samples, one per thread (M <= NTX)
    if (tid < M) nx = load_logical(p, tail, chunk, e_base + (long long)M * (t_first - 1) + 1 + tid);
    float2 cprev[8];
    float pw[8];
#pragma unroll
    for (int r = 0; r < 8; r++) {
        cprev[r] = make_float2(0.f, 0.f);
        pw[r] = 0.f;
    }
    float* dcap = p.d + ((size_t)cap * p.d_rows + DHIST) * N;

    for (int t = t_first - 1; t < t_last; t++) {
        __syncthreads();                                          // ring holds the window of t; fa / fb are free
        const long long e = e_base + (long long)M * t;            // > LE by construction (HTX >= LE + M - 1)
        const long long n_m = (long long)M * ((long long)p.m0 + t) + (M - 1);
        const int nm_mod = (int)(((n_m % N) + N) % N);
        // branches: v[r] = sum_p heq[r + N p] * x[n_m - r - N p], stored at q = (r - n_m) mod N
#pragma unroll
        for (int b = 0; b < 3; b++) {
            const int r = tid + NTX * b;
            const unsigned J = (unsigned)(e - r);
            const unsigned row0 = J / N, col = J - row0 * N;
            const float2 acc = branch_sum_dyn(row0, sm.ring + col, tap[b]);
            int q = r - nm_mod;
            if (q < 0) q += N;
            sm.fa[q] = acc;
        }
        __syncthreads();                                          // window consumed, fa complete
        if (tid < M && t + 1 < t_last) sm.ring[(int)((e + 1 + tid) % RING)] = nx;     // slots older than the next window
        if (tid < M && t + 2 < t_last) nx = load_logical(p, tail, chunk, e + M + 1 + tid);
        // radix-3 pass (NS = 1): 512 butterflies, one per thread
        {
            const int j = tid;
            const float2 a = sm.fa[j], bb = sm.fa[j + N / 3], c = sm.fa[j + 2 * (N / 3)];
            const float s3 = 0.86602540378443865f;
            const float2 bc = cadd(bb, c), dd = csub(bb, c);
            const float2 mid = make_float2(a.x - 0.5f * bc.x, a.y - 0.5f * bc.y);
            const float2 rot = make_float2(-s3 * dd.y, s3 * dd.x);          // i * s3 * (b - c)
            sm.fb[3 * j] = cadd(a, bc);
            sm.fb[3 * j + 1] = cadd(mid, rot);                             // a + b w + c w^2, w = exp(+2 pi i / 3)
            sm.fb[3 * j + 2] = csub(mid, rot);
        }
        __syncthreads();
        if (tid < N / 8) pass8x<3>(sm.fb, sm.fa, sm.tw3[tid % 3], tid);
        __syncthreads();
        if (tid < N / 8) pass8x<24>(sm.fa, sm.fb, sm.tw24[tid % 24], tid);
        __syncthreads();
        if (tid < N / 8) {
            // last pass (NS = 192) in registers: thread j ends up with channels k = j + 192 r, the same at every time
            const int j = tid;
            float2 v[8], w[8];
#pragma unroll
            for (int r = 0; r < 8; r++) v[r] = sm.fb[j + r * (N / 8)];
            powers7(sm.tw192[j], w);
#pragma unroll
            for (int r = 1; r < 8; r++) v[r] = cmul(v[r], w[r]);
            dft8(v);
            if (t >= t_first) {
                float* drow = dcap + (size_t)t * N;
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    const float2 cur = v[r], prv = cprev[r];
                    const float re = cur.x * prv.x + cur.y * prv.y;
                    const float im = cur.y * prv.x - cur.x * prv.y;
                    drow[j + r * (N / 8)] = atan2_branchfree(im, re) * P25_FM_GAIN;
                    pw[r] += cur.x * cur.x + cur.y * cur.y;
                }
                if (p.y) {
                    float2* yrow = p.y + ((size_t)cap * p.y_rows + t) * N;
#pragma unroll
                    for (int r = 0; r < 8; r++) yrow[j + r * (N / 8)] = v[r];
                }
            }
#pragma unroll
            for (int r = 0; r < 8; r++) cprev[r] = v[r];
        }
    }
    if (p.power_sum && tid < N / 8) {
#pragma unroll
        for (int r = 0; r < 8; r++) atomicAdd(p.power_sum + (size_t)cap * N + tid + r * (N / 8), pw[r]);
    }
}

// ------------------------------------------------------------------------------------------------ kernel B
// Boxcar over ten discriminator samples (oldest first, like the oracle) and transpose: a CTA takes 32 channels x 128
// output times of d (128-byte row segments), every warp then writes 32 consecutive times of one channel's baseband row.
constexpr int KC = 32;                 // channels per CTA
constexpr int TB = 128;                // output times per CTA

struct SmemB {
    float dt[TB + DHIST][KC + 1];
};

struct ChanParams {
    const float* d;            // [captures][d_rows][N]; row DHIST + t = output time t of this chunk, rows 0 .. DHIST-1 carried
    float* bb;                 // baseband rows, stream = capture * N + channel
    size_t row_stride;
    unsigned n_out, n_captures, d_rows;
};

__global__ void __launch_bounds__(256) p25_chan_box_kernel(const ChanParams p) {
    __shared__ SmemB sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned cap = blockIdx.z, k0 = blockIdx.y * KC;
    const int t0 = blockIdx.x * TB;
    const float* dc = p.d + (size_t)cap * p.d_rows * N;
    for (int i = warp; i < TB + DHIST; i += 8) {
        const int t = t0 - DHIST + i;                              // row DHIST + t
        sm.dt[i][lane] = t < (int)p.n_out ? __ldg(dc + (size_t)(DHIST + t) * N + k0 + lane) : 0.f;
    }
    __syncthreads();
    for (int ch = warp; ch < KC; ch += 8) {
        float* out = p.bb + ((size_t)cap * N + k0 + ch) * p.row_stride + P25CU_BB_HIST;
        for (int o = lane; o < TB; o += 32) {
            const int t = t0 + o;
            if (t >= (int)p.n_out) break;
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < P25_BOXCAR; i++) acc += sm.dt[o + i][ch];
            out[t] = acc * (1.0f / P25_BOXCAR);
        }
    }
}

// carried input tail: tail_out[i] = logical[n + i], logical = tail_in ++ chunk (both cf32), i < HTX
__global__ void p25_pfb_tail_kernel(const float2* iq, const float2* tail_in, float2* tail_out, unsigned n) {
    const unsigned cap = blockIdx.y;
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (unsigned)HTX) return;
    const unsigned long long l = (unsigned long long)n + i;
    tail_out[(size_t)cap * HTX + i] = l < (unsigned long long)HTX ? tail_in[(size_t)cap * HTX + l] : iq[(size_t)cap * n + (l - HTX)];
}

}  // namespace pfb

// ------------------------------------------------------------------------------------------------ host side
unsigned p25cu_pfb_tail_len() { return (unsigned)pfb::HTX; }
unsigned p25cu_pfb_channels() { return (unsigned)pfb::N; }
unsigned p25cu_pfb_decimation() { return (unsigned)pfb::M; }
unsigned p25cu_pfb_hist_rows() { return (unsigned)pfb::DHIST; }

// Equivalent prototype heq = hp (*) upsample(hc, M), in double, rounded once to f32; twiddle table.
cudaError_t p25cu_pfb_upload(float** d_taps, float2** d_twiddle) {
    cudaError_t e;
    double* acc = new double[pfb::LE]();
    for (int j = 0; j < P25_TAPS_CHAN; j++)
        for (int i = 0; i < pfb::L; i++) acc[i + pfb::M * j] += (double)P25_TAPS_CHAN_H[j] * (double)P25_TAPS_PFB_H[i];
    float* heq = new float[pfb::LE];
    for (int i = 0; i < pfb::LE; i++) heq[i] = (float)acc[i];
    delete[] acc;
    if ((e = cudaMalloc(d_taps, sizeof(float) * pfb::LE)) == cudaSuccess)
        e = cudaMemcpy(*d_taps, heq, sizeof(float) * pfb::LE, cudaMemcpyHostToDevice);
    delete[] heq;
    if (e != cudaSuccess) return e;
    float2* tw = new float2[pfb::N];
    for (int t = 0; t < pfb::N; t++) {
        const double a = 2.0 * 3.14159265358979323846 * (double)t / (double)pfb::N;
        tw[t] = make_float2((float)cos(a), (float)sin(a));
    }
    if ((e = cudaMalloc(d_twiddle, sizeof(float2) * pfb::N)) == cudaSuccess)
        e = cudaMemcpy(*d_twiddle, tw, sizeof(float2) * pfb::N, cudaMemcpyHostToDevice);
    delete[] tw;
    return e;
}

// Per-device setup (once per device under the library's plan mutex, that device current).
cudaError_t p25cu_pfb_plan_device(P25DevPlan* plan) {
    cudaError_t e = cudaFuncSetAttribute(pfb::p25_pfbx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(pfb::SmemX));
    if (e != cudaSuccess) return e;
    plan->pfb_slots = plan->n_sm;          // one 222 KB CTA per SM
    return cudaSuccess;
}

// One chunk of every capture: discriminator rows into d (rows DHIST ..), baseband rows into bb, new tail.
cudaError_t p25cu_launch_pfb(const void* iq, const void* tail_in, void* tail_out, const float* taps, const float2* twiddle, float* d,
                             unsigned d_rows, float2* y, unsigned y_rows, float* bb, size_t row_stride, float* power_sum,
                             unsigned long long a0, unsigned long long m0, unsigned n, unsigned n_out, unsigned n_captures,
                             cudaStream_t st, unsigned* launches, const P25DevPlan* plan) {
    if (n_out) {
        pfb::PfbParams a;
        a.iq = (const float2*)iq;
        a.tail_in = (const float2*)tail_in;
        a.taps = taps;
        a.twiddle = twiddle;
        a.d = d;
        a.y = y;
        a.power_sum = power_sum;
        a.a0 = a0;
        a.m0 = m0;
        a.n = n;
        a.n_out = n_out;
        a.n_captures = n_captures;
        a.d_rows = d_rows;
        a.y_rows = y_rows;
        // one wave: every SM gets one run of consecutive output times (each run pays one window fill and one warm-up
        // output time), the captures share the SMs evenly
        unsigned per_cap = (unsigned)plan->pfb_slots / n_captures;
        if (per_cap < 1) per_cap = 1;
        unsigned tpc = (n_out + per_cap - 1) / per_cap;                  // output times per CTA
        if (tpc < 8) tpc = n_out < 8 ? n_out : 8;
        const dim3 ga((n_out + tpc - 1) / tpc, n_captures);
        pfb::p25_pfbx_kernel<<<ga, pfb::NTX, sizeof(pfb::SmemX), st>>>(a, tpc);
        pfb::ChanParams b;
        b.d = d;
        b.bb = bb;
        b.row_stride = row_stride;
        b.n_out = n_out;
        b.n_captures = n_captures;
        b.d_rows = d_rows;
        const dim3 gb((n_out + pfb::TB - 1) / pfb::TB, pfb::N / pfb::KC, n_captures);
        pfb::p25_chan_box_kernel<<<gb, 256, 0, st>>>(b);
        *launches += 2;
    }
    if (n) {
        const dim3 gt((pfb::HTX + 255) / 256, n_captures);
        pfb::p25_pfb_tail_kernel<<<gt, 256, 0, st>>>((const float2*)iq, (const float2*)tail_in, (float2*)tail_out, n);
        *launches += 1;
    }
    return cudaGetLastError();
}
