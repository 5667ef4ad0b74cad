"""Dataflow model of the integer-tensor-pipe /5 decimator (p25rx_b200/csrc/ddc_fm.cu, `w5i::p25_ddc5_imma_kernel`).

Not the oracle and not product code: this restates in numpy, lane for lane, what the kernel asks of
mma.sync.m16n8k32.s32.u8.s8 -- the fragment layouts of the PTX ISA, the banded tap matrix in fragment order
(`w5i::Tables`), the permuted slice rows, the 24-bit taps in three balanced limbs, the accumulator start values and the
magic-number conversion -- so that tests/test_w5i_dataflow.py can check the decomposition on the CPU against the
float64 FIR before any GPU run.

m16n8k32 fragments (lane = 4 g + t): A regs a0..a3 hold row g / g+8, k bytes 4t..4t+3 and 16+4t..16+4t+3;
B regs b0, b1 hold column g, the same k bytes; C regs c0..c3 hold (row g, cols 2t, 2t+1), (row g+8, cols 2t, 2t+1).
"""
from __future__ import annotations

import numpy as np

NTAPS = 25
SCALE_LOG2 = 25
R, NOUT, XNEW = 8, 256, 1280
MAGIC = 0x4B400000          # bit pattern of 1.5 * 2^23


def limbs(taps: np.ndarray):
    """24-bit fixed-point taps as three balanced s8 digits; returns (limb[k][l], sum of the quantised taps)."""
    out = np.zeros((len(taps), 3), dtype=np.int64)
    tsum = 0
    for k, h in enumerate(taps):
        t = int(np.round(float(h) * (1 << SCALE_LOG2)))
        v = t
        for l in range(3):
            d = ((v + 128) & 255) - 128
            out[k, l] = d
            v = (v - d) // 256
        assert v == 0, "tap does not fit 24 bits"
        tsum += t
    return out, tsum


def b_table(lb: np.ndarray, sk: int) -> np.ndarray:
    """B fragments [nt][jj][limb][lane][reg][byte] as signed bytes (the kernel's g_btab[sk])."""
    tab = np.zeros((2, 3, 3, 32, 2, 4), dtype=np.int64)
    for nt in range(2):
        for jj in range(3):
            ks = nt + jj
            for l in range(3):
                for lane in range(32):
                    n, t = lane >> 2, lane & 3
                    for h in range(2):
                        for bb in range(4):
                            phi = 32 * ks + 16 * h + 4 * t + bb          # byte of the row window
                            smp, comp = phi >> 1, phi & 1
                            d = smp - sk - 5 * (4 * nt + (n >> 1))
                            if comp == (n & 1) and 0 <= d < NTAPS:
                                tab[nt, jj, l, lane, h, bb] = lb[NTAPS - 1 - d, l]
    return tab


def mma_u8s8(acc, a_frag, b_frag):
    """acc[lane][4] += A x B for one warp.  a_frag[lane][4 regs][4 bytes] (u8), b_frag[lane][2 regs][4 bytes] (s8)."""
    A = np.zeros((16, 32), dtype=np.int64)
    B = np.zeros((32, 8), dtype=np.int64)
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for bb in range(4):
            A[g, 4 * t + bb] = a_frag[lane, 0, bb]
            A[g + 8, 4 * t + bb] = a_frag[lane, 1, bb]
            A[g, 16 + 4 * t + bb] = a_frag[lane, 2, bb]
            A[g + 8, 16 + 4 * t + bb] = a_frag[lane, 3, bb]
            B[4 * t + bb, g] = b_frag[lane, 0, bb]
            B[16 + 4 * t + bb, g] = b_frag[lane, 1, bb]
    D = A @ B
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        acc[lane, 0] += D[g, 2 * t]
        acc[lane, 1] += D[g, 2 * t + 1]
        acc[lane, 2] += D[g + 8, 2 * t]
        acc[lane, 3] += D[g + 8, 2 * t + 1]


def decimate_iteration(xs: np.ndarray, skew: int, taps: np.ndarray):
    """One warp iteration.  xs: the staged slice as bytes (I, Q interleaved) from its 16-byte aligned start; the
    iteration's first input sample sits `skew` samples (0..7) into it.  Returns (y[256] complex, in byte units x 2^25,
    as float32 pairs exactly like the kernel's FADD2 / FFMA2 sequence, bank lists of the loads and stores)."""
    lb, _ = limbs(taps)
    tab = b_table(lb, skew & 3)
    init = [MAGIC - 128 * int(lb[:, l].sum()) for l in range(3)]
    acc = np.zeros((2, 2, 3, 32, 4), dtype=np.int64)
    acc += np.array(init, dtype=np.int64)[None, None, :, None, None]
    lane = np.arange(32)
    g, t = lane >> 2, lane & 3
    drow = ((g & 3) << 1) | (g >> 2)
    a_off = 80 * drow + 8 * (skew >> 2) + 4 * t
    load_banks = []
    for ks in range(4):
        for mt in range(2):
            ad = a_off + 1280 * mt + 32 * ks
            a_frag = np.zeros((32, 4, 4), dtype=np.int64)
            for r, off in enumerate((0, 640, 16, 656)):
                for ln in range(32):
                    a_frag[ln, r] = xs[ad[ln] + off: ad[ln] + off + 4]
                load_banks.append(((ad + off) // 4) % 32)
            for nt in range(2):
                jj = ks - nt
                if jj < 0 or jj > 2:
                    continue
                for l in range(3):
                    mma_u8s8(acc[mt, nt, l], a_frag, tab[nt, jj, l])
    y = np.zeros(NOUT, dtype=np.complex128)
    yf = np.zeros((NOUT, 2), dtype=np.float32)
    store_words = []
    for mt in range(2):
        for nt in range(2):
            for h in range(2):
                words = []
                for ln in range(32):
                    o = 128 * mt + 8 * (drow[ln] + 8 * h) + 4 * nt + t[ln]
                    f = []
                    for l in range(3):
                        bits = acc[mt, nt, l, ln, 2 * h: 2 * h + 2]
                        assert np.all(np.abs(bits - MAGIC) < (1 << 22)), "accumulator leaves the magic number's binade"
                        f.append((np.array(bits, dtype=np.uint32).view(np.float32) - np.float32(12582912.0)).astype(np.float32))
                    inner = (np.float64(f[1]) * 256.0 + np.float64(f[0])).astype(np.float32)     # one rounding, like FFMA
                    outer = (np.float64(f[2]) * 65536.0 + np.float64(inner)).astype(np.float32)
                    yf[o] = outer
                    # shared-memory word the lane's 8-byte store starts at: array q, row, half
                    q, half = 2 * nt + (t[ln] >> 1), t[ln] & 1
                    row = 5 + 16 * mt + drow[ln] + 8 * h
                    words.append(q * 37 * 4 + row * 4 + 2 * half)
                store_words.append(np.array(words))
    return yf, load_banks, store_words


# ---------------------------------------------------------------------------------------------------------------------
# u8 /50 (`w50i::p25_ddc50_imma_kernel`): the /10 front stage and the /5 decimator as ONE 290-tap FIR on the raw bytes.
# Rows are 200 samples (400 bytes) apart and hold four outputs (one n-tile), 28 k-steps x 3 limbs per m-tile.
# ---------------------------------------------------------------------------------------------------------------------
D50, G50, KS50, SCALE50 = 50, 290, 28, 28


def combined_taps(hf: np.ndarray, hd: np.ndarray) -> np.ndarray:
    """g[10 k + i] = hd[k] hf[i] in float64 from the float32 tap sets (yd[t] = sum_j g[j] X[50 t + 49 - j])."""
    g = np.zeros(len(hf) + 10 * (len(hd) - 1))
    for k, h in enumerate(hd.astype(np.float64)):
        g[10 * k: 10 * k + len(hf)] += h * hf.astype(np.float64)
    return g


def limbs50(g: np.ndarray):
    out = np.zeros((len(g), 3), dtype=np.int64)
    for j, h in enumerate(g):
        v = int(np.round(float(h) * (1 << SCALE50)))
        for l in range(3):
            d = ((v + 128) & 255) - 128
            out[j, l] = d
            v = (v - d) // 256
        assert v == 0, "tap does not fit 24 bits"
    return out


def b_table50(lb: np.ndarray, sk: int) -> np.ndarray:
    tab = np.zeros((KS50, 3, 32, 2, 4), dtype=np.int64)
    for ks in range(KS50):
        for l in range(3):
            for lane in range(32):
                n, t = lane >> 2, lane & 3
                for h in range(2):
                    for bb in range(4):
                        phi = 32 * ks + 16 * h + 4 * t + bb
                        smp, comp = phi >> 1, phi & 1
                        d = smp - sk - D50 * (n >> 1)
                        if comp == (n & 1) and 0 <= d < G50:
                            tab[ks, l, lane, h, bb] = lb[G50 - 1 - d, l]
    return tab


def decimate50_iteration(xs: np.ndarray, skew: int, g: np.ndarray):
    """One warp iteration (128 outputs).  xs: staged slice bytes from its 16-byte aligned start; the first sample of the
    first output's 290-sample window sits `skew` samples (0..7) into it.  Returns (y[128][2] float32, load banks, stores)."""
    lb = limbs50(g)
    tab = b_table50(lb, skew & 1)
    assert all(128 * int(np.abs(lb[:, l]).sum()) < (1 << 22) for l in range(3)), "accumulators stay in the magic binade"
    init = [MAGIC - 128 * int(lb[:, l].sum()) for l in range(3)]
    acc = np.zeros((2, 3, 32, 4), dtype=np.int64) + np.array(init, dtype=np.int64)[None, :, None, None]
    lane = np.arange(32)
    gq, t = lane >> 2, lane & 3
    a_off = 400 * gq + 4 * (skew >> 1) + 4 * t
    load_banks = []
    for ks in range(KS50):
        for mt in range(2):
            ad = a_off + 6400 * mt + 32 * ks
            a_frag = np.zeros((32, 4, 4), dtype=np.int64)
            for r, off in enumerate((0, 3200, 16, 3216)):
                for ln in range(32):
                    a_frag[ln, r] = xs[ad[ln] + off: ad[ln] + off + 4]
                load_banks.append(((ad + off) // 4) % 32)
            for l in range(3):
                mma_u8s8(acc[mt, l], a_frag, tab[ks, l])
    yf = np.zeros((128, 2), dtype=np.float32)
    stores = []
    for mt in range(2):
        for h in range(2):
            words = []
            for ln in range(32):
                o = 4 * (16 * mt + gq[ln] + 8 * h) + t[ln]
                f = []
                for l in range(3):
                    bits = acc[mt, l, ln, 2 * h: 2 * h + 2]
                    assert np.all(np.abs(bits - MAGIC) < (1 << 22))
                    f.append((np.array(bits, dtype=np.uint32).view(np.float32) - np.float32(12582912.0)).astype(np.float32))
                inner = (np.float64(f[1]) * 256.0 + np.float64(f[0])).astype(np.float32)
                yf[o] = (np.float64(f[2]) * 65536.0 + np.float64(inner)).astype(np.float32)
                arr, half = t[ln] >> 1, t[ln] & 1                     # ydA (44 rows) | ydB
                words.append(arr * 44 * 4 + (10 + 16 * mt + gq[ln] + 8 * h) * 4 + 2 * half)
            stores.append(np.array(words))
    return yf, load_banks, stores
