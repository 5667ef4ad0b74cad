// pipe_peaks.cu -- measured instruction-pipe and copy ceilings of the GPU this runs on (B200, sm_100a).
//
// The /5 demod kernels and the channelizer are bound by the FP32 pipe, not by HBM; MEASURED_PEAKS.json
// only holds an HBM copy figure and a cuBLAS bf16 figure, so the roofline of those kernels needs its own measured
// denominators.  This program times dependent-chain-free loops of
//   FFMA (3-register), FFMA2 (fma.rn.f32x2), mma.sync m16n8k8 tf32, mma.sync m16n8k16 bf16, IDP4A
// with 2,048 resident threads per SM, and a pinned host -> device copy (the ceiling of the end-to-end leg of bench.py).
// Output: one JSON object on stdout.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_peaks tools/pipe_peaks.cu && ./pipe_peaks
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int ITERS = 4096;
constexpr int CHAINS = 8;

__global__ void __launch_bounds__(256) k_ffma(float* out, float a, float b) {
    float acc[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) acc[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s += acc[i];
    if (s == 12345.678f) out[0] = s;
}

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

__global__ void __launch_bounds__(256) k_ffma2(float* out, float a, float b) {
    unsigned long long acc[CHAINS];
    const float2 aa = make_float2(a, a), bb = make_float2(b, b);
    const unsigned long long A = *reinterpret_cast<const unsigned long long*>(&aa), B = *reinterpret_cast<const unsigned long long*>(&bb);
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
        const float2 v = make_float2(threadIdx.x * 1e-3f + i, 1.f);
        acc[i] = *reinterpret_cast<const unsigned long long*>(&v);
    }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) acc[i] = fma2(acc[i], A, B);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
        const float2 v = *reinterpret_cast<float2*>(&acc[i]);
        s += v.x + v.y;
    }
    if (s == 12345.678f) out[0] = s;
}

// FFMA2 and FFMA interleaved 1:1 (do they share one pipe?)
__global__ void __launch_bounds__(256) k_mix(float* out, float a, float b) {
    unsigned long long acc2[CHAINS / 2];
    float acc[CHAINS / 2];
    const float2 aa = make_float2(a, a), bb = make_float2(b, b);
    const unsigned long long A = *reinterpret_cast<const unsigned long long*>(&aa), B = *reinterpret_cast<const unsigned long long*>(&bb);
#pragma unroll
    for (int i = 0; i < CHAINS / 2; i++) {
        const float2 v = make_float2(threadIdx.x * 1e-3f + i, 1.f);
        acc2[i] = *reinterpret_cast<const unsigned long long*>(&v);
        acc[i] = v.x;
    }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS / 2; i++) {
            acc2[i] = fma2(acc2[i], A, B);
            acc[i] = fmaf(acc[i], a, b);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CHAINS / 2; i++) {
        const float2 v = *reinterpret_cast<float2*>(&acc2[i]);
        s += v.x + v.y + acc[i];
    }
    if (s == 12345.678f) out[0] = s;
}

__global__ void __launch_bounds__(256) k_mma_tf32(float* out, unsigned a, unsigned b) {
    float d[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) d[i][j] = 0.f;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                         : "r"(a), "r"(a + 1), "r"(a + 2), "r"(a + 3), "r"(b), "r"(b + 1));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) s += d[i][j];
    if (s == 12345.678f) out[0] = s;
}

__global__ void __launch_bounds__(256) k_mma_bf16(float* out, unsigned a, unsigned b) {
    float d[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) d[i][j] = 0.f;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                         : "r"(a), "r"(a + 1), "r"(a + 2), "r"(a + 3), "r"(b), "r"(b + 1));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) s += d[i][j];
    if (s == 12345.678f) out[0] = s;
}

// tensor pipe beside the FP32 pipe: one mma.sync tf32 per 8 FFMA2 (do the two pipes run concurrently?)
__global__ void __launch_bounds__(256) k_mma_plus_ffma2(float* out, unsigned a, unsigned b, float fa, float fb) {
    float d[2][4];
    unsigned long long acc[CHAINS];
    const float2 aa = make_float2(fa, fa), bb = make_float2(fb, fb);
    const unsigned long long A = *reinterpret_cast<const unsigned long long*>(&aa), B = *reinterpret_cast<const unsigned long long*>(&bb);
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) d[i][j] = 0.f;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
        const float2 v = make_float2(threadIdx.x * 1e-3f + i, 1.f);
        acc[i] = *reinterpret_cast<const unsigned long long*>(&v);
    }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 2; i++) {
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                         : "r"(a), "r"(a + 1), "r"(a + 2), "r"(a + 3), "r"(b), "r"(b + 1));
#pragma unroll
            for (int j = 0; j < CHAINS; j++) acc[j] = fma2(acc[j], A, B);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) s += d[i][j];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
        const float2 v = *reinterpret_cast<float2*>(&acc[i]);
        s += v.x + v.y;
    }
    if (s == 12345.678f) out[0] = s;
}

__global__ void __launch_bounds__(256) k_dp4a(int* out, int a, int b) {
    int acc[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) acc[i] = __dp4a(a + acc[i], b, acc[i]);
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s += acc[i];
    if (s == 123456789) out[0] = s;
}

// PRMT (alu pipe) interleaved with FFMA2 (fma pipe): the u8 -> f32 conversion of the /5 kernel costs issue slots only if the pipes overlap
__global__ void __launch_bounds__(256) k_prmt_plus_ffma2(float* out, unsigned w, float fa, float fb) {
    unsigned long long acc[CHAINS];
    unsigned p[CHAINS];
    const float2 aa = make_float2(fa, fa), bb = make_float2(fb, fb);
    const unsigned long long A = *reinterpret_cast<const unsigned long long*>(&aa), B = *reinterpret_cast<const unsigned long long*>(&bb);
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
        const float2 v = make_float2(threadIdx.x * 1e-3f + i, 1.f);
        acc[i] = *reinterpret_cast<const unsigned long long*>(&v);
        p[i] = w + i + threadIdx.x * 0x01010101u;
    }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            acc[i] = fma2(acc[i], A, B);
            p[i] = __byte_perm(p[i], 0x4B000000u, 0x7440u + (p[i] & 1));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
        const float2 v = *reinterpret_cast<float2*>(&acc[i]);
        s += v.x + v.y + __uint_as_float(p[i]);
    }
    if (s == 12345.678f) out[0] = s;
}

// legacy integer tensor pipe: mma.sync m16n8k32 u8 x s8 -> s32 (the /5 decimator on raw IQ bytes)
__global__ void __launch_bounds__(256) k_imma(int* out, unsigned a, unsigned b) {
    int d[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) d[i][j] = 0;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(d[i][0]), "+r"(d[i][1]), "+r"(d[i][2]), "+r"(d[i][3])
                         : "r"(a), "r"(a + 1), "r"(a + 2), "r"(a + 3), "r"(b), "r"(b + 1));
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) s += d[i][j];
    if (s == 123456789) out[0] = s;
}

// the mix of the IMMA decimator kernel: 1 IMMA per 12 FFMA2 (two IMMA chains, 24 FFMA2 per round)
__global__ void __launch_bounds__(256) k_imma_plus_ffma2(float* out, unsigned a, unsigned b, float fa, float fb) {
    int d[2][4];
    unsigned long long acc[CHAINS];
    const float2 aa = make_float2(fa, fa), bb = make_float2(fb, fb);
    const unsigned long long A = *reinterpret_cast<const unsigned long long*>(&aa), B = *reinterpret_cast<const unsigned long long*>(&bb);
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) d[i][j] = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
        const float2 v = make_float2(threadIdx.x * 1e-3f + i, 1.f);
        acc[i] = *reinterpret_cast<const unsigned long long*>(&v);
    }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 2; i++)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(d[i][0]), "+r"(d[i][1]), "+r"(d[i][2]), "+r"(d[i][3])
                         : "r"(a), "r"(a + 1), "r"(a + 2), "r"(a + 3), "r"(b), "r"(b + 1));
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int i = 0; i < CHAINS; i++) acc[i] = fma2(acc[i], A, B);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) s += (float)d[i][j];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
        const float2 v = *reinterpret_cast<float2*>(&acc[i]);
        s += v.x + v.y;
    }
    if (s == 12345.678f) out[0] = s;
}

// int -> float conversions (I2F) per clock: one way to turn the IMMA accumulators into floats
__global__ void __launch_bounds__(256) k_i2f(float* out, int a) {
    int v[CHAINS];
    float f[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) { v[i] = a + threadIdx.x + i; f[i] = 0.f; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            f[i] = __int2float_rn(v[i]);
            v[i] = __float_as_int(f[i]) ^ it;
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s += f[i];
    if (s == 12345.678f) out[0] = s;
}

template <typename F>
static double time_ms(F launch, int reps = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    int dev = 0, n_sm = 0, clk = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev));
    float* d_out;
    CK(cudaMalloc(&d_out, 64));
    const int grid = n_sm * 8, block = 256;
    const double threads = (double)grid * block;
    const double t_ffma = time_ms([&] { k_ffma<<<grid, block>>>(d_out, 1.0001f, 0.5f); });
    const double t_ffma2 = time_ms([&] { k_ffma2<<<grid, block>>>(d_out, 1.0001f, 0.5f); });
    const double t_mix = time_ms([&] { k_mix<<<grid, block>>>(d_out, 1.0001f, 0.5f); });
    const double t_tf32 = time_ms([&] { k_mma_tf32<<<grid, block>>>(d_out, 0x3f800000u, 0x3f000000u); });
    const double t_bf16 = time_ms([&] { k_mma_bf16<<<grid, block>>>(d_out, 0x3f803f80u, 0x3f003f00u); });
    const double t_both = time_ms([&] { k_mma_plus_ffma2<<<grid, block>>>(d_out, 0x3f800000u, 0x3f000000u, 1.0001f, 0.5f); });
    const double t_dp4a = time_ms([&] { k_dp4a<<<grid, block>>>((int*)d_out, 0x01020304, 0x01010101); });
    const double t_prmt = time_ms([&] { k_prmt_plus_ffma2<<<grid, block>>>(d_out, 0x12345678u, 1.0001f, 0.5f); });
    const double t_imma = time_ms([&] { k_imma<<<grid, block>>>((int*)d_out, 0x01020304u, 0x01010101u); });
    const double t_imma_both = time_ms([&] { k_imma_plus_ffma2<<<grid, block>>>(d_out, 0x01020304u, 0x01010101u, 1.0001f, 0.5f); });
    const double t_i2f = time_ms([&] { k_i2f<<<grid, block>>>(d_out, 12345); });
    CK(cudaGetLastError());
    const double fma_ffma = threads * ITERS * CHAINS / (t_ffma * 1e-3);            // FMA / s
    const double fma_ffma2 = threads * ITERS * CHAINS * 2 / (t_ffma2 * 1e-3);
    const double fma_mix = threads * ITERS * (CHAINS / 2) * 3 / (t_mix * 1e-3);
    const double warps = threads / 32;
    const double mac_tf32 = warps * ITERS * 4 * (16.0 * 8 * 8) / (t_tf32 * 1e-3);  // MAC / s
    const double mac_bf16 = warps * ITERS * 4 * (16.0 * 8 * 16) / (t_bf16 * 1e-3);
    const double both_mac = warps * ITERS * 2 * (16.0 * 8 * 8) / (t_both * 1e-3);
    const double both_fma = threads * ITERS * 2 * CHAINS * 2 / (t_both * 1e-3);
    const double dp4a = threads * ITERS * CHAINS * 4 / (t_dp4a * 1e-3);
    const double prmt_fma = threads * ITERS * CHAINS * 2 / (t_prmt * 1e-3);
    const double mac_imma = warps * ITERS * 4 * (16.0 * 8 * 32) / (t_imma * 1e-3);
    const double imma_both_mac = warps * ITERS * 2 * (16.0 * 8 * 32) / (t_imma_both * 1e-3);
    const double imma_both_fma = threads * ITERS * 3 * CHAINS * 2 / (t_imma_both * 1e-3);
    const double i2f_rate = threads * ITERS * CHAINS / (t_i2f * 1e-3);

    // pinned host -> device copy ceiling (what bounds bench.py's e2e leg), 1 GiB x 5, best
    size_t bytes = (size_t)1 << 30;
    void *h = nullptr, *d = nullptr;
    double h2d = 0.0, d2h = 0.0;
    if (cudaHostAlloc(&h, bytes, cudaHostAllocDefault) == cudaSuccess && cudaMalloc(&d, bytes) == cudaSuccess) {
        memset(h, 1, bytes);
        const double t = time_ms([&] { cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, 0); });
        h2d = bytes / (t * 1e-3) / 1e9;
        const double t2 = time_ms([&] { cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, 0); });
        d2h = bytes / (t2 * 1e-3) / 1e9;
    }
    printf("{\"n_sm\": %d, \"sm_clock_khz_max\": %d, "
           "\"fp32_ffma_tflops\": %.2f, \"fp32_ffma2_tflops\": %.2f, \"fp32_ffma_ffma2_mix_tflops\": %.2f, "
           "\"fp32_fma_per_clk_per_sm_ffma\": %.1f, \"fp32_fma_per_clk_per_sm_ffma2\": %.1f, "
           "\"mma_sync_tf32_tflops\": %.1f, \"mma_sync_bf16_tflops\": %.1f, "
           "\"mma_sync_tf32_mac_per_clk_per_sm\": %.1f, "
           "\"concurrent_tf32_tflops\": %.1f, \"concurrent_ffma2_tflops\": %.2f, "
           "\"dp4a_tops\": %.1f, \"ffma2_tflops_with_1_prmt_each\": %.2f, "
           "\"mma_sync_u8s8_k32_tops\": %.1f, \"mma_sync_u8s8_instr_clk_per_subcore\": %.2f, "
           "\"concurrent_imma_tops\": %.1f, \"concurrent_imma_ffma2_tflops\": %.2f, \"i2f_per_clk_per_sm\": %.1f, "
           "\"h2d_pinned_gbs\": %.1f, \"d2h_pinned_gbs\": %.1f, "
           "\"how\": \"%d CTAs x 256 threads, %d iterations x %d independent chains per thread, best of 5, CUDA events; flops = 2 x FMA\"}\n",
           n_sm, clk, 2 * fma_ffma / 1e12, 2 * fma_ffma2 / 1e12, 2 * fma_mix / 1e12, fma_ffma / n_sm / (clk * 1e3),
           fma_ffma2 / n_sm / (clk * 1e3), 2 * mac_tf32 / 1e12, 2 * mac_bf16 / 1e12, mac_tf32 / n_sm / (clk * 1e3),
           2 * both_mac / 1e12, 2 * both_fma / 1e12, 2 * dp4a / 1e12, 2 * prmt_fma / 1e12,
           2 * mac_imma / 1e12, (clk * 1e3) * n_sm * 4 / (mac_imma / (16.0 * 8 * 32)), 2 * imma_both_mac / 1e12, 2 * imma_both_fma / 1e12,
           i2f_rate / n_sm / (clk * 1e3), h2d, d2h, grid, ITERS, CHAINS);
    return 0;
}
