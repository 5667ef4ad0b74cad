"""CPU restatement of the wideband channelizer (float64 numpy).

ORACLE -- TEST INFRASTRUCTURE ONLY.  The reference (kchmck/p25rx) has no channelizer: one RTL-SDR tuner, one
channel, retuned by hopping (reference src/sdr.rs:61-68, src/recv.rs:127-137); BASELINE.json configs[2] asks for
one, so its behaviour is defined by spec/p25_spec.py (PFB_* constants, taps_pfb) and restated here twice:

  channel_direct  the definition: mix down by k * 12.5 kHz, prototype low-pass, keep every 400th output
  channelize      the same numbers through the polyphase identity and numpy's FFT

tests/test_pfb_oracle.py checks the two against each other; the GPU kernels (p25rx_b200/csrc/pfb.cu) are
checked against channelize().  Everything after the channelizer is the reference's own 48 kHz chain
(oracle.pyoracle.DemodChain(..., front=2)).
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "spec"))
import p25_spec as S  # noqa: E402

N, M, P = S.PFB_CHANNELS, S.PFB_DECIM, S.PFB_TAPS_PER_BRANCH
L = N * P


def _taps() -> np.ndarray:
    return S.taps_pfb().astype(np.float64)


def channel_direct(x: np.ndarray, k: int, m: np.ndarray) -> np.ndarray:
    """y_k[m] = sum_i h[i] x[n_m - i] exp(-2j pi k (n_m - i) / N), n_m = M m + M - 1; x[n] = 0 for n < 0."""
    h = _taps()
    x = np.asarray(x, dtype=np.complex128)
    out = np.zeros(len(m), dtype=np.complex128)
    i = np.arange(L)
    for j, mm in enumerate(np.asarray(m)):
        idx = M * int(mm) + M - 1 - i
        ok = idx >= 0
        xs = np.zeros(L, dtype=np.complex128)
        xs[ok] = x[idx[ok]]
        ph = np.exp(-2j * np.pi * ((k * idx) % N) / N)
        out[j] = np.sum(h * xs * ph)
    return out


def channelize(x: np.ndarray, m0: int = 0, n_out: int | None = None) -> np.ndarray:
    """All channels for output times m0 .. m0 + n_out - 1: returns [n_out][N] complex128.
    x holds the stream from absolute sample 0 (zeros before it)."""
    h = _taps().reshape(P, N)                       # h[p][r] = h[r + N p]
    x = np.asarray(x, dtype=np.complex128)
    if n_out is None:
        n_out = len(x) // M - m0
    xp = np.concatenate([np.zeros(L, dtype=np.complex128), x])   # xp[L + n] = x[n]
    out = np.empty((n_out, N), dtype=np.complex128)
    r = np.arange(N)
    for j in range(n_out):
        nm = M * (m0 + j) + M - 1
        win = xp[L + nm - L + 1: L + nm + 1][::-1]              # win[i] = x[nm - i], i = 0 .. L-1
        v = np.sum(h * win.reshape(P, N), axis=0)               # v[r] = sum_p h[r + N p] x[nm - r - N p]
        u = v[(r + nm) % N]
        out[j] = np.fft.ifft(u) * N                             # sum_q u[q] exp(+2j pi k q / N)
    return out


def channel_filter(y: np.ndarray) -> np.ndarray:
    """The reference's channel-select FIR (src/demod.rs:93) down every channel column of y [n_out][N], zeros before the
    stream start: c_k[m] = sum_j hc[j] y_k[m - j].  This is what the CUDA kernels deliver as channel spectra: they fold
    the filter into the polyphase prototype (p25rx_b200/csrc/pfb.cu)."""
    hc = S.taps_chan().astype(np.float64)
    y = np.asarray(y, dtype=np.complex128)
    out = np.zeros_like(y)
    for j, h in enumerate(hc[: len(y)]):
        out[j:] += h * y[: len(y) - j]
    return out


def equivalent_prototype() -> np.ndarray:
    """heq = hp (*) upsample(hc, M): the prototype whose polyphase channelizer output equals channel_filter(channelize(x))."""
    hp, hc = _taps(), S.taps_chan().astype(np.float64)
    heq = np.zeros(L + M * (len(hc) - 1), dtype=np.float64)
    for j, h in enumerate(hc):
        heq[M * j: M * j + L] += h * hp
    return heq


def channelize_fused(x: np.ndarray, m0: int = 0, n_out: int | None = None) -> np.ndarray:
    """channel_filter(channelize(x)) computed the way the kernels do: one polyphase pass with the equivalent prototype
    (15 taps per branch), float32-rounded taps like the device table."""
    heq = equivalent_prototype()
    PE = -(-len(heq) // N)
    h = np.zeros(PE * N)
    h[: len(heq)] = heq.astype(np.float32).astype(np.float64)
    h = h.reshape(PE, N)
    LE = PE * N
    x = np.asarray(x, dtype=np.complex128)
    if n_out is None:
        n_out = len(x) // M - m0
    xp = np.concatenate([np.zeros(LE, dtype=np.complex128), x])
    out = np.empty((n_out, N), dtype=np.complex128)
    r = np.arange(N)
    for j in range(n_out):
        nm = M * (m0 + j) + M - 1
        win = xp[nm + 1: LE + nm + 1][::-1]                     # win[i] = x[nm - i], i = 0 .. LE-1
        v = np.sum(h * win.reshape(PE, N), axis=0)
        out[j] = np.fft.ifft(v[(r + nm) % N]) * N
    return out
