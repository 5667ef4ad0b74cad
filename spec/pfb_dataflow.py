"""Dataflow model of the round-2 channelizer kernel (p25rx_b200/csrc/pfb.cu, `p25_pfbc_kernel`), in numpy.

Not the oracle (that is oracle/pfb_oracle.py) and not product code: this restates, thread for thread, the index
arithmetic of the kernel -- class-stationary register windows that advance a warp at a time, the tap rows indexed by
the distance back to the newest sample (early samples meet a zero tap), the parity split over the two CTAs of a
cluster, the 8 x 8 x 12 Stockham passes with the skewed buffer and the radix-2 combine -- so that
tests/test_pfb_oracle.py can check the decomposition against the float64 definition on the CPU, before any GPU run.

Definitions (pfb.cu header):  c_k[m] = sum_i heq[i] x[n_m - i] exp(-2j pi k (n_m - i) / N),  n_m = M m + M - 1,
i = r + N p, q = (r - n_m) mod N:  c_k[m] = sum_q V[q] exp(+2j pi k q / N),  V[q] = sum_p heq[r + N p] x[n_m - r - N p].
Every sample contributing to V[q] has index s = -q (mod N): FFT input q belongs to the residue class c = (-q) mod N.
"""
from __future__ import annotations

import numpy as np

N, M, PE = 1536, 400, 15
LE = N * PE
H = N // 2                 # FFT length per parity
NTH = 768                  # threads per CTA: one class (FFT input u = tid) and one channel each
TB = 8                     # output times per batch
WARM = 10                  # warm-up output times per run (1 for c[m-1], 9 for the boxcar history)
WS = 16                    # window slots per class
EARLY = 64                 # a window's newest sample lies at most this far ahead of n_m
ROWS = (N + EARLY) // 2
HTX = 27648                # carried input tail: >= LE + (WARM + 1) * M + EARLY, a multiple of 128


def skew_a(i):
    """Position of element i in the buffer between passes A and B (one pad per 16)."""
    return i + (i >> 4)


def skew_b(i):
    """Position of element i in the buffer between passes B and C (8 pads per 64)."""
    return i + 8 * (i >> 6)


def tap_table(heq: np.ndarray, b: int) -> np.ndarray:
    """Rows of CTA b: tab[row][k] = heq[d0 + N k] (0 outside 0 .. LE-1), d0 = 2 row - EARLY + (1 - b)."""
    h = np.zeros(LE, dtype=np.float32)
    h[: len(heq)] = heq.astype(np.float32)
    d0 = 2 * np.arange(ROWS) - EARLY + (1 - b)
    tab = np.zeros((ROWS, WS), dtype=np.float32)
    for k in range(WS):
        idx = d0 + N * k
        ok = (idx >= 0) & (idx < LE)
        tab[ok, k] = h[idx[ok]]
    return tab


def stockham_pass(buf_in, R, NS, n, sign=+1):
    """One autosort pass on a length-n vector: out[(j // NS) NS R + j % NS + q NS] = DFT_R(in[j + r n / R] w^(k r))."""
    out = np.zeros(n, dtype=buf_in.dtype)
    j = np.arange(n // R)
    k = j % NS
    v = np.stack([buf_in[j + r * (n // R)] * np.exp(sign * 2j * np.pi * k * r / (NS * R)) for r in range(R)])
    for q in range(R):
        w = sum(v[r] * np.exp(sign * 2j * np.pi * r * q / R) for r in range(R))
        out[(j // NS) * NS * R + k + q * NS] = w
    return out


def idft768(v):
    a = stockham_pass(v, 8, 1, H)
    b = stockham_pass(a, 8, 8, H)
    return stockham_pass(b, 12, 64, H)


class Run:
    """One cluster: output times [t_first, t_last) of one capture, relative to the chunk start."""

    def __init__(self, heq, tail, chunk, a0, m0, dtype=np.complex128):
        self.tabs = [tap_table(heq, b).astype(np.float64) for b in (0, 1)]
        self.logical = np.concatenate([tail, chunk]).astype(dtype)      # index l = s - a0 + HTX
        self.a0, self.m0 = a0, m0
        self.e_base = M * m0 + (M - 1) - a0 + HTX

    def load(self, l):
        l = np.asarray(l)
        ok = (l >= 0) & (l < len(self.logical))
        out = np.zeros(l.shape, dtype=self.logical.dtype)
        out[ok] = self.logical[l[ok]]
        return out

    def spectra(self, t_first, t_last):
        """[t_last - t_first + WARM][N] channel spectra c_k for times t_first - WARM .. t_last - 1."""
        a0, e_base = self.a0, self.e_base
        t_init = t_first - WARM - 1                                     # state describes this time; every later time advances first
        e_init = e_base + M * t_init
        st = []
        for b in (0, 1):
            u = np.arange(NTH)                                          # FFT input of thread tid: u = tid
            lane = u % 32
            c = (-(2 * u + b)) % N
            cl = (c - a0 + HTX) % N
            assert e_init >= N
            rp = (e_init - cl) % N
            rp0 = rp[u - lane]                                          # lane 0 of the same warp
            d0 = rp0 + 2 * lane - np.where(rp0 + 62 >= N, N, 0)         # the warp's windows are 32 consecutive samples
            assert np.all((d0 - rp) % N == 0) and np.all(d0 >= -EARLY) and np.all(d0 < N)
            lnew = e_init - d0
            slots = np.zeros((WS, len(u)), dtype=self.logical.dtype)
            for k in range(WS):
                slots[(-k) % WS] = self.load(lnew - N * k)
            st.append({"d0": d0, "phi": np.zeros(len(u), dtype=np.int64), "lim": N - M - 2 * (31 - lane), "slots": slots})
        out = []
        for t in range(t_first - WARM, t_last):
            e = e_base + M * t
            EO = []
            for b in (0, 1):
                s = st[b]
                adv = s["d0"] >= s["lim"]
                assert all(len(set(adv[w:w + 32])) == 1 for w in range(0, len(adv), 32))      # warp-uniform
                s["d0"] = s["d0"] + np.where(adv, M - N, M)
                assert np.all(s["d0"] >= -EARLY) and np.all(s["d0"] < N) and np.all((s["d0"] + EARLY) % 2 == 1 - b)
                s["phi"] = (s["phi"] + adv) % WS
                nx = self.load(e - s["d0"])
                idx = np.nonzero(adv)[0]
                s["slots"][s["phi"][idx], idx] = nx[idx]
                row = (s["d0"] + EARLY) >> 1
                V = np.zeros(H, dtype=self.logical.dtype)
                for j in range(WS):
                    V += s["slots"][j] * self.tabs[b][row, (s["phi"] - j) % WS]
                EO.append(idft768(V))                                   # thread order == FFT input order (u = tid)
            E, O = EO
            k = np.arange(H)
            w = np.exp(2j * np.pi * k / N)
            out.append(np.concatenate([E + w * O, E - w * O]))
        return np.array(out)
