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
// Kernel A (p25_pfb_kernel): a CTA owns a run of consecutive output times; for each it forms the branches from the
// input window in global memory (consecutive windows overlap by 93 %: L1/L2 hits) and runs the DFT as a mixed-radix
// Stockham FFT 3 x 8 x 8 x 8 in shared memory (radix-8 butterflies in registers, twiddles from a 1,536-entry table);
// the last pass writes the spectrum time-major, Y[m][k], fully coalesced.
// Kernel B (p25_chan_fm_kernel): a CTA takes 32 channels x 128 output times of Y (256-byte row segments, lanes =
// channels), runs the channel filter as a sliding register window down each lane's column, the discriminator
// and the boxcar, and transposes through shared memory so that every channel's baseband row is written in
// 128-byte segments.  HBM-bound: 8 B in + 2 x 30.7 B (Y out, Y in) + 15.4 B out per input sample.
#include "p25cu_internal.cuh"
#include "p25_pfb_taps.h"

namespace pfb {

constexpr int N = P25_PFB_N, M = P25_PFB_M, P = P25_PFB_P, L = N * P;
constexpr int NT = 256;
static_assert(N == 3 * 8 * 8 * 8, "FFT plan is 3 x 8 x 8 x 8");

// Shared memory holds only the FFT ping-pong buffers and the twiddles (37 KB -> 6 CTAs per SM).  The input window
// (6,144 samples per output time, 93 % of it shared with the previous output time of the same CTA) is read straight
// from global memory: the re-reads hit L1/L2, and the occupancy this buys matters more than the staging it saves
// (staged window: 2 CTAs per SM, 23 % issue utilisation).
struct SmemA {
    float2 fa[N], fb[N];
    float2 tw[N];
};

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

// one radix-8 Stockham pass: thread j of 192, NS = product of the radices already done
template <int NS, bool LAST>
__device__ __forceinline__ void pass8(const float2* __restrict__ in, float2* __restrict__ out, const float2* __restrict__ tw, int j) {
    const int k = j % NS;
    float2 v[8];
#pragma unroll
    for (int r = 0; r < 8; r++) v[r] = in[j + r * (N / 8)];
#pragma unroll
    for (int r = 1; r < 8; r++) v[r] = cmul(v[r], tw[r * k * (N / (NS * 8))]);
    dft8(v);
    const int j0 = (j / NS) * NS * 8 + k;
#pragma unroll
    for (int r = 0; r < 8; r++) out[j0 + r * NS] = v[r];
}

struct PfbParams {
    const float2* iq;          // [captures][n]
    const float2* tail_in;     // [captures][L]
    const float* taps;         // [L] prototype
    const float2* twiddle;     // [N] exp(+2 pi i t / N)
    float2* y;                 // [captures][y_rows][N], this chunk's rows start at row hist
    unsigned long long a0, m0; // absolute input / output index of the chunk start
    unsigned n, n_out, n_captures;
    unsigned y_rows, hist;
};

__global__ void __launch_bounds__(NT) p25_pfb_kernel(const PfbParams p, const unsigned times_per_cta) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemA& sm = *reinterpret_cast<SmemA*>(smem_raw);
    const int tid = threadIdx.x;
    const unsigned cap = blockIdx.y;
    const unsigned t_first = blockIdx.x * times_per_cta;         // this CTA's consecutive output times, relative to m0
    const unsigned t_last = min(t_first + times_per_cta, p.n_out);
    const float2* chunk = p.iq + (size_t)cap * p.n;
    const float2* tail = p.tail_in + (size_t)cap * L;
    for (int i = tid; i < N; i += NT) sm.tw[i] = p.twiddle[i];
    float2* yc = p.y + ((size_t)cap * p.y_rows + p.hist) * N;
    __syncthreads();

    for (unsigned tr = t_first; tr < t_last; tr++) {
        // newest input of output m (absolute) is n_m = M m + M - 1; its logical index in (tail ++ chunk) is e
        const long long n_m = (long long)M * ((long long)p.m0 + tr) + (M - 1);
        const long long e = n_m - (long long)p.a0 + L;           // tail holds L samples
        const int nm_mod = (int)((unsigned long long)n_m % N);
        const bool in_chunk = e - (L - 1) >= L;                  // the whole window lies in this chunk
        // branches: v[r] = sum_p h[r + N p] * x[n_m - r - N p], stored at q = (r - n_m) mod N
#pragma unroll
        for (int jr = 0; jr < N / NT; jr++) {
            const int r = tid + jr * NT;
            float2 acc = make_float2(0.f, 0.f);
            if (in_chunk) {
                const float2* x = chunk + (e - L - r);
#pragma unroll
                for (int pp = 0; pp < P; pp++) acc = cfma(__ldg(p.taps + r + N * pp), __ldg(x - N * pp), acc);
            } else {
#pragma unroll
                for (int pp = 0; pp < P; pp++) {
                    const long long l = e - r - N * pp;
                    float2 xv = make_float2(0.f, 0.f);
                    if (l >= L) xv = __ldg(chunk + (l - L));
                    else if (l >= 0) xv = tail[l];
                    acc = cfma(__ldg(p.taps + r + N * pp), xv, acc);
                }
            }
            int q = r - nm_mod;
            if (q < 0) q += N;
            sm.fa[q] = acc;
        }
        __syncthreads();
        // radix-3 pass (NS = 1): 512 butterflies
        for (int j = tid; j < N / 3; j += NT) {
            const float2 a = sm.fa[j], b = sm.fa[j + N / 3], c = sm.fa[j + 2 * (N / 3)];
            const float s3 = 0.86602540378443865f;
            const float2 bc = cadd(b, c), d = csub(b, c);
            const float2 mid = make_float2(a.x - 0.5f * bc.x, a.y - 0.5f * bc.y);
            const float2 rot = make_float2(-s3 * d.y, s3 * d.x);          // i * s3 * (b - c)
            sm.fb[3 * j] = cadd(a, bc);
            sm.fb[3 * j + 1] = cadd(mid, rot);                             // a + b w + c w^2, w = exp(+2 pi i / 3)
            sm.fb[3 * j + 2] = csub(mid, rot);
        }
        __syncthreads();
        if (tid < N / 8) pass8<3, false>(sm.fb, sm.fa, sm.tw, tid);
        __syncthreads();
        if (tid < N / 8) pass8<24, false>(sm.fa, sm.fb, sm.tw, tid);
        __syncthreads();
        if (tid < N / 8) pass8<192, true>(sm.fb, yc + (size_t)tr * N, sm.tw, tid);   // natural order, straight to HBM
        __syncthreads();                                         // fb is read until here; fa is rewritten next
    }
}

// ------------------------------------------------------------------------------------------------ kernel B
constexpr int KC = 32;                 // channels per CTA
constexpr int TB = 128;                // output times per CTA
constexpr int HC = P25_TAPS_CHAN - 1;  // 40
constexpr int HALO = HC + 1 + (P25_BOXCAR - 1);   // 50
constexpr int RC = 6;

struct SmemB {
    float2 yt[TB + HALO][KC];
    float2 ct[TB + 10 + RC][KC];
    float dt[TB + 9][KC + 1];
    float pw[KC];
};

__constant__ float c_chan[P25_TAPS_CHAN];

struct ChanParams {
    const float2* y;           // [captures][y_rows][N]
    float* bb;                 // baseband rows, stream = capture * N + channel
    size_t row_stride;
    float* power_sum;          // [streams] or null
    unsigned n_out, n_captures, y_rows, hist;
};

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

__global__ void __launch_bounds__(256) p25_chan_fm_kernel(const ChanParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmemB& sm = *reinterpret_cast<SmemB*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned cap = blockIdx.z, k0 = blockIdx.y * KC;
    const int t0 = blockIdx.x * TB;                               // first output time of the tile, relative to m0
    const float2* yc = p.y + (size_t)cap * p.y_rows * N;
    // rows of Y: row hist + t holds output time t of this chunk; rows [0, hist) are the previous chunk's last ones
    for (int i = warp; i < TB + HALO; i += 8) {
        const int t = t0 - HALO + i;
        float2 v = make_float2(0.f, 0.f);
        if (t < (int)p.n_out) v = __ldg(yc + (size_t)((int)p.hist + t) * N + k0 + lane);
        sm.yt[i][lane] = v;
    }
    if (tid < KC) sm.pw[tid] = 0.f;
    __syncthreads();
    // channel-select FIR down each lane's column: c at time t0 - 10 + tc uses yt rows tc .. tc + 40
    for (int blk = warp; blk * RC < TB + 10; blk += 8) {
        const int tc0 = blk * RC;
        float2 acc[RC];
#pragma unroll
        for (int r = 0; r < RC; r++) acc[r] = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < HC + RC; i++) {
            const int row = tc0 + i;
            const float2 x = row < TB + HALO ? sm.yt[row][lane] : make_float2(0.f, 0.f);
#pragma unroll
            for (int r = 0; r < RC; r++) {
                const int k = HC - i + r;
                if (k >= 0 && k <= HC) acc[r] = cfma(c_chan[k], x, acc[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < RC; r++) sm.ct[tc0 + r][lane] = acc[r];
    }
    __syncthreads();
    // discriminator: d at time t0 - 9 + td from c rows td + 1, td; power of the stored outputs' c
    float pw = 0.f;
    for (int td = warp; td < TB + 9; td += 8) {
        const float2 prv = sm.ct[td][lane], cur = sm.ct[td + 1][lane];
        const float re = cur.x * prv.x + cur.y * prv.y;
        const float im = cur.y * prv.x - cur.x * prv.y;
        sm.dt[td][lane] = atan2_branchfree(im, re) * P25_FM_GAIN;
        const int t = t0 - 9 + td;
        if (t >= t0 && t < (int)p.n_out) pw += cur.x * cur.x + cur.y * cur.y;
    }
    if (p.power_sum) atomicAdd(&sm.pw[lane], pw);
    __syncthreads();
    // boxcar + transpose: a warp writes 32 consecutive output times of one channel
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
    if (p.power_sum && tid < KC) atomicAdd(p.power_sum + (size_t)cap * N + k0 + tid, sm.pw[tid]);
}

// carried input tail: tail_out[i] = logical[n + i], logical = tail_in ++ chunk (both cf32), i < L
__global__ void p25_pfb_tail_kernel(const float2* iq, const float2* tail_in, float2* tail_out, unsigned n) {
    const unsigned cap = blockIdx.y;
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (unsigned)L) return;
    const unsigned long long l = (unsigned long long)n + i;
    tail_out[(size_t)cap * L + i] = l < (unsigned long long)L ? tail_in[(size_t)cap * L + l] : iq[(size_t)cap * n + (l - L)];
}

}  // namespace pfb

// ------------------------------------------------------------------------------------------------ host side
unsigned p25cu_pfb_tail_len() { return (unsigned)pfb::L; }
unsigned p25cu_pfb_channels() { return (unsigned)pfb::N; }
unsigned p25cu_pfb_decimation() { return (unsigned)pfb::M; }
unsigned p25cu_pfb_hist_rows() { return 64; }

cudaError_t p25cu_pfb_upload(float** d_taps, float2** d_twiddle) {
    cudaError_t e;
    if ((e = cudaMemcpyToSymbol(pfb::c_chan, P25_TAPS_CHAN_H, sizeof(float) * P25_TAPS_CHAN)) != cudaSuccess) return e;
    if ((e = cudaMalloc(d_taps, sizeof(float) * pfb::L)) != cudaSuccess) return e;
    if ((e = cudaMemcpy(*d_taps, P25_TAPS_PFB_H, sizeof(float) * pfb::L, cudaMemcpyHostToDevice)) != cudaSuccess) return e;
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
    cudaError_t e = cudaFuncSetAttribute(pfb::p25_pfb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(pfb::SmemA));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(pfb::p25_chan_fm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(pfb::SmemB));
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pfb::p25_pfb_kernel, pfb::NT, sizeof(pfb::SmemA));
    if (e != cudaSuccess) return e;
    plan->pfb_slots = plan->n_sm * (per_sm > 0 ? per_sm : 1);
    return cudaSuccess;
}

// One chunk of every capture: spectrum rows into y (rows hist ..), baseband rows into bb, new tail.
cudaError_t p25cu_launch_pfb(const void* iq, const void* tail_in, void* tail_out, const float* taps, const float2* twiddle, float2* y,
                             unsigned y_rows, float* bb, size_t row_stride, float* power_sum, unsigned long long a0,
                             unsigned long long m0, unsigned n, unsigned n_out, unsigned n_captures, cudaStream_t st,
                             unsigned* launches, const P25DevPlan* plan) {
    const unsigned hist = p25cu_pfb_hist_rows();
    if (n_out) {
        pfb::PfbParams a;
        a.iq = (const float2*)iq;
        a.tail_in = (const float2*)tail_in;
        a.taps = taps;
        a.twiddle = twiddle;
        a.y = y;
        a.a0 = a0;
        a.m0 = m0;
        a.n = n;
        a.n_out = n_out;
        a.n_captures = n_captures;
        a.y_rows = y_rows;
        a.hist = hist;
        // one wave of CTAs: consecutive output times per CTA (their windows overlap, so re-reads hit L1), as many CTAs
        // as fit at once
        const int slots = plan->pfb_slots;
        unsigned per_cap = (unsigned)slots / n_captures;
        if (per_cap < 1) per_cap = 1;
        const unsigned tpc = (n_out + per_cap - 1) / per_cap;            // output times per CTA
        const dim3 ga((n_out + tpc - 1) / tpc, n_captures);
        pfb::p25_pfb_kernel<<<ga, pfb::NT, sizeof(pfb::SmemA), st>>>(a, tpc);
        pfb::ChanParams b;
        b.y = y;
        b.bb = bb;
        b.row_stride = row_stride;
        b.power_sum = power_sum;
        b.n_out = n_out;
        b.n_captures = n_captures;
        b.y_rows = y_rows;
        b.hist = hist;
        const dim3 gb((n_out + pfb::TB - 1) / pfb::TB, pfb::N / pfb::KC, n_captures);
        pfb::p25_chan_fm_kernel<<<gb, 256, sizeof(pfb::SmemB), st>>>(b);
        *launches += 2;
    }
    if (n) {
        const dim3 gt((pfb::L + 255) / 256, n_captures);
        pfb::p25_pfb_tail_kernel<<<gt, 256, 0, st>>>((const float2*)iq, (const float2*)tail_in, (float2*)tail_out, n);
        *launches += 1;
    }
    return cudaGetLastError();
}
