// p25_imma_tables.h -- host-side tables of the integer-tensor-pipe decimators (ddc_fm.cu: w5i, w50i).  Plain C++, no CUDA
// types: included by ddc_fm.cu, which uploads them, and by tests/hostcheck/imma_hostcheck.cpp, which hands the very same
// tables to the CPU suite (tests/test_w5i_dataflow.py compares them with the numpy model spec/w5i_dataflow.py).
//
// A FIR y[o] = sum_j t[j] x[D o + (T - 1) - j] on interleaved u8 I/Q bytes as mma.sync.m16n8k32.s32.u8.s8:
//   A = rows of the staged slice (raw bytes), B = the taps as a banded matrix -- column 2 n' + c (output n' of the row,
//   component c) carries tap T - 1 - d at the byte of sample sk + D n + d, component c -- in three balanced s8 limbs of
//   t = round(h 2^S).  Fragment order (PTX ISA, lane = 4 g + tig): b0 / b1 of column g hold k bytes 4 tig .. + 3 and
//   16 + 4 tig .. + 3 of the k-step.  The accumulators start at the bit pattern of 1.5 * 2^23 minus 128 x (limb sum), so they
//   are the floats 1.5 * 2^23 + sum l (u - 128) as long as |sum| < 2^22 (`ok`).
#pragma once
#include <math.h>

#include "p25_tables.h"

namespace p25imma {

struct Frag {
    unsigned x, y;      // b0, b1 (layout of uint2)
};

constexpr int MAGIC = 0x4B400000;       // bit pattern of 1.5 * 2^23

// balanced base-256 digits of t: t = d0 + 256 d1 + 65536 d2, each in [-128, 127]; false if t needs more than 24 bits
inline bool split_limbs(long long t, int (&d)[3]) {
    long long v = t;
    for (int l = 0; l < 3; l++) {
        const long long q = ((v + 128) & 255) - 128;
        d[l] = (int)q;
        v = (v - q) / 256;
    }
    return v == 0;
}

// ---- /5 decimator on 240 kS/s u8 (w5i): 25 taps, rows of eight outputs = two n-tiles x three k-steps, four skews
constexpr int NB5 = 18, SCALE5 = 25;
struct Tables5 {
    Frag b[4][NB5][32];                 // [skew & 3][(n-tile * 3 + k-step - n-tile) * 3 + limb][lane]
    int init[3];
    double gain;                        // sum of the quantised taps / 2^S: DC gain of the integer filter
    bool ok;
    explicit Tables5(const float* taps) {
        static_assert(P25_TAPS_DECIM == 25 && P25_DECIM_NATIVE == 5, "geometry");
        int limb[P25_TAPS_DECIM][3];
        long long sum[3] = {0, 0, 0}, asum[3] = {0, 0, 0}, tsum = 0;
        ok = true;
        for (int k = 0; k < P25_TAPS_DECIM; k++) {
            const long long t = llround((double)taps[k] * (double)(1 << SCALE5));
            if (!split_limbs(t, limb[k])) ok = false;
            for (int l = 0; l < 3; l++) {
                sum[l] += limb[k][l];
                asum[l] += limb[k][l] < 0 ? -limb[k][l] : limb[k][l];
            }
            tsum += t;
        }
        for (int l = 0; l < 3; l++)
            if (128 * asum[l] >= (1 << 22)) ok = false;
        gain = (double)tsum / (double)(1 << SCALE5);
        for (int l = 0; l < 3; l++) init[l] = MAGIC - 128 * (int)sum[l];
        for (int sk = 0; sk < 4; sk++)
            for (int nt = 0; nt < 2; nt++)
                for (int jj = 0; jj < 3; jj++)
                    for (int l = 0; l < 3; l++)
                        for (int ln = 0; ln < 32; ln++) {
                            const int n = ln >> 2, tig = ln & 3, ks = nt + jj;
                            unsigned w[2] = {0u, 0u};
                            for (int h = 0; h < 2; h++)
                                for (int bb = 0; bb < 4; bb++) {
                                    const int phi = 32 * ks + 16 * h + 4 * tig + bb;    // byte of the row window
                                    const int smp = phi >> 1, comp = phi & 1;
                                    const int d = smp - sk - 5 * (4 * nt + (n >> 1));
                                    int v = 0;
                                    if (comp == (n & 1) && d >= 0 && d < P25_TAPS_DECIM) v = limb[P25_TAPS_DECIM - 1 - d][l];
                                    w[h] |= (unsigned)(v & 255) << (8 * bb);
                                }
                            b[sk][(nt * 3 + jj) * 3 + l][ln] = Frag{w[0], w[1]};
                        }
    }
};

// ---- /50 on 2.4 MS/s u8 (w50i): front /10 (50 taps) and decimator /5 (25 taps) as one FIR, g[10 k + i] = hd[k] hf[i]; rows
// of four outputs = one n-tile x 28 k-steps, two skews
constexpr int G50 = P25_TAPS_FRONT + P25_DECIM_FRONT * (P25_TAPS_DECIM - 1);   // 290
constexpr int KS50 = 28, NB50 = KS50 * 3, SCALE50 = 28;
struct Tables50 {
    Frag b[2][NB50][32];                // [skew & 1][k-step * 3 + limb][lane]
    int init[3];
    double gain;
    bool ok;
    Tables50(const float* hf, const float* hd) {
        double gd[G50];
        for (int j = 0; j < G50; j++) gd[j] = 0.0;
        for (int k = 0; k < P25_TAPS_DECIM; k++)
            for (int i = 0; i < P25_TAPS_FRONT; i++) gd[P25_DECIM_FRONT * k + i] += (double)hd[k] * (double)hf[i];
        static int limb[G50][3];
        long long sum[3] = {0, 0, 0}, asum[3] = {0, 0, 0}, tsum = 0;
        ok = true;
        for (int j = 0; j < G50; j++) {
            const long long t = llround(gd[j] * (double)(1 << SCALE50));
            if (!split_limbs(t, limb[j])) ok = false;                  // a tap that does not fit 24 bits
            for (int l = 0; l < 3; l++) {
                sum[l] += limb[j][l];
                asum[l] += limb[j][l] < 0 ? -limb[j][l] : limb[j][l];
            }
            tsum += t;
        }
        for (int l = 0; l < 3; l++)
            if (128 * asum[l] >= (1 << 22)) ok = false;                // an accumulator could leave the binade of 1.5 * 2^23
        gain = (double)tsum / (double)(1 << SCALE50);
        for (int l = 0; l < 3; l++) init[l] = MAGIC - 128 * (int)sum[l];
        for (int sk = 0; sk < 2; sk++)
            for (int ks = 0; ks < KS50; ks++)
                for (int l = 0; l < 3; l++)
                    for (int ln = 0; ln < 32; ln++) {
                        const int n = ln >> 2, tig = ln & 3;
                        unsigned w[2] = {0u, 0u};
                        for (int h = 0; h < 2; h++)
                            for (int bb = 0; bb < 4; bb++) {
                                const int phi = 32 * ks + 16 * h + 4 * tig + bb;        // byte of the row window
                                const int smp = phi >> 1, comp = phi & 1;
                                const int d = smp - sk - 50 * (n >> 1);                 // y[t] = sum_j g[j] X[50 t + 289 - j]
                                int v = 0;
                                if (comp == (n & 1) && d >= 0 && d < G50) v = limb[G50 - 1 - d][l];
                                w[h] |= (unsigned)(v & 255) << (8 * bb);
                            }
                        b[sk][ks * 3 + l][ln] = Frag{w[0], w[1]};
                    }
    }
};

}  // namespace p25imma
