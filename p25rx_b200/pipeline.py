"""Host-side mirror of the reference's hot-path interfaces over the C ABI.

  DemodTask        <-> reference src/demod.rs:25-119   (one run-loop iteration per call)
  MessageReceiver  <-> p25::message::receiver::MessageReceiver as used at src/recv.rs:81,:207,:136
  ReplayReceiver   <-> reference src/replay.rs:11-57
  power_dbm        <-> reference src/demod.rs:123-134 (computed inside DemodTask.run_chunk)

Every class is batched over `n_streams` independent streams that live on one GPU.  All
arithmetic happens in libp25cu.so (CUDA, sm_100a); this module only moves pointers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import EVENT_DTYPE, FMT_CF32_IQ, FMT_U8_IQ, P25Error

# MessageEvent variants (reference src/recv.rs:214-233)
EV_ERROR, EV_NID, EV_VOICE_HEADER, EV_LINK_CONTROL, EV_CRYPTO_CONTROL, EV_LSD, EV_VOICE_FRAME, EV_TSBK, EV_VOICE_TERM = range(9)
EVENT_NAMES = ["Error", "PacketNID", "VoiceHeader", "LinkControl", "CryptoControl", "LowSpeedDataFragment",
               "VoiceFrame", "TrunkingControl", "VoiceTerm"]
STATS_FAMILIES = ["bch", "cyclic", "golayStd", "golayExt", "golayShort", "hammingStd", "hammingShort",
                  "rsShort", "rsMed", "rsLong", "viterbiDibit", "viterbiTribit"]   # reference src/hub.rs:557-572

BUF_BYTES = 32768            # reference src/consts.rs:6
BUF_SAMPLES = BUF_BYTES // 2  # reference src/consts.rs:8


def _as_ptr(x):
    """numpy array, torch tensor (host or device) or int address -> (void*, on_device)."""
    if isinstance(x, np.ndarray):
        return x.ctypes.data_as(C.c_void_p), False
    if hasattr(x, "data_ptr"):  # torch.Tensor without importing torch here
        return C.c_void_p(x.data_ptr()), bool(x.is_cuda)
    raise TypeError(f"unsupported buffer type {type(x)}")


class Context:
    """One p25cu_ctx: n_streams streams on one GPU."""

    def __init__(self, n_streams: int, fmt: int = FMT_U8_IQ, decimation: int = 5, max_chunk_samples: int = BUF_SAMPLES,
                 max_baseband: int = 0, device: int = 0, event_slots: int = 0):
        self._L = _lib.lib()
        cfg = _lib.Config(device, n_streams, fmt, decimation, max_chunk_samples, max_baseband, _lib.ABI_VERSION, event_slots)
        h = C.c_void_p()
        rc = self._L.p25cu_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise P25Error(rc, self._L.p25cu_last_error(None).decode())
        self._h = h
        self.n_streams, self.fmt, self.decimation = n_streams, fmt, decimation
        self.max_chunk_samples = max_chunk_samples
        self._a_abs = 0  # input samples consumed per stream (decimator phase bookkeeping)

    def close(self):
        if getattr(self, "_h", None):
            self._L.p25cu_destroy(self._h)
            self._h = None

    __del__ = close

    def _ck(self, rc: int, ok=(0,)):
        if rc not in ok:
            raise P25Error(rc, self._L.p25cu_last_error(self._h).decode())
        return rc

    # ---- Surface 1
    def demod(self, iq, n_in_per_stream: int, want_baseband: bool = True, want_power: bool = False):
        """p25cu_demod.  iq: numpy array (host) or torch tensor (host or cuda), [S][n] samples of the
        context's format ([S / 1536][n] wideband captures when decimation = 400).
        Returns (baseband [S][n_out] or None, n_out, power_dbm [S] or None)."""
        ptr, on_dev = _as_ptr(iq)
        n = int(n_in_per_stream)
        n_expect = (self._a_abs + n) // self.decimation - self._a_abs // self.decimation
        bb = np.empty((self.n_streams, n_expect), dtype=np.float32) if want_baseband else None
        pw = np.empty(self.n_streams, dtype=np.float32) if want_power else None
        n_out = C.c_size_t(0)
        self._ck(self._L.p25cu_demod(self._h, ptr, n, int(on_dev), bb.ctypes.data_as(C.c_void_p) if want_baseband else None,
                                     C.byref(n_out), pw.ctypes.data_as(C.c_void_p) if want_power else None))
        self._a_abs += n
        assert n_out.value == n_expect
        return bb, n_out.value, pw

    # ---- Surface 2
    def decode(self, baseband: np.ndarray | None = None):
        if baseband is None:
            self._ck(self._L.p25cu_decode(self._h, None, 0))
            return
        bb = np.ascontiguousarray(baseband, dtype=np.float32).reshape(self.n_streams, -1)
        self._ck(self._L.p25cu_decode(self._h, bb.ctypes.data_as(C.c_void_p), bb.shape[1]))

    def process(self, iq, n_in_per_stream: int):
        ptr, on_dev = _as_ptr(iq)
        self._ck(self._L.p25cu_process(self._h, ptr, n_in_per_stream, int(on_dev)))
        self._a_abs += int(n_in_per_stream)

    def pending(self) -> int:
        n = C.c_size_t(0)
        self._ck(self._L.p25cu_pending(self._h, C.byref(n)))
        return n.value

    def poll(self, cap: int | None = None, copy: bool = True) -> np.ndarray:
        """Drain queued events ordered by (stream, sample).  copy=False returns a view of the context's
        pinned buffer (p25cu_poll_view), valid until the next poll."""
        if cap is None:
            ptr, n = C.c_void_p(), C.c_size_t(0)
            self._ck(self._L.p25cu_poll_view(self._h, C.byref(ptr), C.byref(n)))
            if n.value == 0:
                return np.zeros(0, dtype=EVENT_DTYPE)
            buf = (C.c_char * (n.value * EVENT_DTYPE.itemsize)).from_address(ptr.value)
            view = np.frombuffer(buf, dtype=EVENT_DTYPE, count=n.value)
            return view.copy() if copy else view
        ev = np.zeros(max(cap, 1), dtype=EVENT_DTYPE)
        n = C.c_size_t(0)
        self._ck(self._L.p25cu_poll(self._h, ev.ctypes.data_as(C.c_void_p), cap, C.byref(n)))
        return ev[: n.value]

    # ---- packed, asynchronous drain (p25cu_poll_start / p25cu_poll_packed)
    def poll_start(self):
        """Queue the compaction of everything decoded so far; returns at once (at most two outstanding)."""
        self._ck(self._L.p25cu_poll_start(self._h))

    def poll_packed(self, copy: bool = True):
        """Collect the oldest started poll (starting one if none is outstanding): (u32 words, n_events, more).
        copy=False returns a view of the context's pinned buffer, valid until two further polls have been started."""
        ptr, nw, ne, more = C.c_void_p(), C.c_size_t(0), C.c_size_t(0), C.c_int(0)
        self._ck(self._L.p25cu_poll_packed(self._h, C.byref(ptr), C.byref(nw), C.byref(ne), C.byref(more)))
        if nw.value == 0:
            return np.zeros(0, dtype=np.uint32), 0, bool(more.value)
        buf = (C.c_uint32 * nw.value).from_address(ptr.value)
        words = np.frombuffer(buf, dtype=np.uint32, count=nw.value)
        return (words.copy() if copy else words), ne.value, bool(more.value)

    def unpack(self, words: np.ndarray, n_events: int) -> np.ndarray:
        """Packed records -> the 80-byte event records poll() returns (host-side byte shuffling)."""
        words = np.ascontiguousarray(words, dtype=np.uint32)
        ev = np.zeros(max(n_events, 1), dtype=EVENT_DTYPE)
        n = C.c_size_t(0)
        self._ck(self._L.p25cu_unpack_events(words.ctypes.data_as(C.c_void_p), words.size, ev.ctypes.data_as(C.c_void_p), n_events,
                                             C.byref(n)))
        return ev[: n.value]

    # ---- pinned host buffers (the reference's buffer pools, src/demod.rs:63, src/sdr.rs:25-33)
    def host_alloc(self, shape, dtype) -> np.ndarray:
        """A page-locked numpy array owned by the library; release with host_free(arr)."""
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dtype.itemsize
        p = C.c_void_p()
        self._ck(self._L.p25cu_host_alloc(self._h, nbytes, C.byref(p)))
        buf = (C.c_char * nbytes).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
        self._pinned = getattr(self, "_pinned", {})
        self._pinned[arr.ctypes.data] = p.value
        return arr

    def host_free(self, arr: np.ndarray):
        p = self._pinned.pop(arr.ctypes.data)
        self._ck(self._L.p25cu_host_free(self._h, C.c_void_p(p)))

    def host_register(self, arr: np.ndarray):
        self._ck(self._L.p25cu_host_register(self._h, arr.ctypes.data_as(C.c_void_p), arr.nbytes))

    def host_unregister(self, arr: np.ndarray):
        self._ck(self._L.p25cu_host_unregister(self._h, arr.ctypes.data_as(C.c_void_p)))

    def resync(self, stream: int):
        self._ck(self._L.p25cu_resync(self._h, stream))

    def stats(self, stream: int, clear: bool = False) -> np.ndarray:
        st = _lib.Stats()
        self._ck(self._L.p25cu_get_stats(self._h, stream, C.byref(st), int(clear)))
        return np.ctypeslib.as_array(st.code).astype(np.uint64).reshape(12, 4).copy()

    def read_baseband(self, stream: int, n: int) -> np.ndarray:
        """First n baseband samples of one stream of the last demod / process (device -> host)."""
        out = np.empty(n, dtype=np.float32)
        self._ck(self._L.p25cu_read_baseband(self._h, stream, out.ctypes.data_as(C.c_void_p), n))
        return out

    def keep_spectra(self, on: bool = True):
        """Channelizer mode test hook: keep the channel-filtered spectra of every following demod for channelizer_output()."""
        self._ck(self._L.p25cu_set_keep_spectra(self._h, int(on)))

    def channelizer_output(self) -> np.ndarray:
        """Channelizer mode (decimation 400): channel spectra (after the channel-select filter) of the last demod,
        [captures][n_out][1536] complex64; needs keep_spectra() before that demod."""
        n = C.c_size_t(0)
        self._ck(self._L.p25cu_channelizer_output(self._h, None, C.byref(n)))
        out = np.zeros((self.n_streams // 1536, n.value, 1536), dtype=np.complex64)
        if n.value:
            self._ck(self._L.p25cu_channelizer_output(self._h, out.ctypes.data_as(C.c_void_p), C.byref(n)))
        return out

    def sync(self):
        self._ck(self._L.p25cu_sync(self._h))

    def set_overlap(self, on: bool):
        self._ck(self._L.p25cu_set_overlap(self._h, int(on)))

    def demod_timing(self, enable: bool):
        """(average demod-kernel ms, launches) recorded since the last call; sets whether recording continues."""
        ms, n = C.c_double(0), C.c_uint(0)
        self._ck(self._L.p25cu_demod_timing(self._h, int(enable), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    @property
    def cuda_stream(self) -> int:
        return int(self._L.p25cu_cuda_stream(self._h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self._L.p25cu_launch_count(self._h))

    def fec_selftest(self, kind: int, words: np.ndarray, n: int = 0, k: int = 0):
        words = np.ascontiguousarray(words)
        warp_kind, kind = kind, {10: 7, 11: 9, 12: 0, 13: 8, 15: 14}.get(kind, kind)   # 10..13, 15: warp-cooperative forms
        if kind == 0:
            count, out = words.size, np.zeros(words.size, dtype=np.uint32)
        elif kind == 7:
            count, out = words.size // n, None
        elif kind == 8:
            count = words.size // 98
            out = np.zeros((count, 12), dtype=np.uint8)
        elif kind == 9:
            count = words.size // 72
            out = np.zeros((count, 15), dtype=np.uint32)
        elif kind == 14:
            count = words.size // 98
            out = np.zeros((count, 18), dtype=np.uint8)
        else:
            count, out = words.size, np.zeros(words.size, dtype=np.uint32)
        nerr = np.zeros(count, dtype=np.int32)
        self._ck(self._L.p25cu_fec_selftest(self._h, warp_kind, words.ctypes.data_as(C.c_void_p), count, n, k,
                                            out.ctypes.data_as(C.c_void_p) if out is not None else None,
                                            nerr.ctypes.data_as(C.c_void_p)))
        return (words if kind == 7 else out), nerr


class DemodTask:
    """Batched DemodTask (reference src/demod.rs:25-119): IQ chunk in, 48 kHz baseband chunk out."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        self._notifier = 0  # Throttler::new(4), reference src/demod.rs:67

    def run_chunk(self, iq, n_in_per_stream: int | None = None):
        """One loop iteration (src/demod.rs:70-117).  Returns (baseband [S][n_out], power_dbm or None);
        like the reference the signal power is reported on every 4th chunk only (src/demod.rs:95-101)."""
        want_power = self._notifier == 0
        self._notifier = (self._notifier + 1) % 4
        if n_in_per_stream is None:
            n_in_per_stream = iq.size // self.ctx.n_streams // (2 if self.ctx.fmt == FMT_U8_IQ else 1)
        bb, _, pw = self.ctx.demod(iq, n_in_per_stream, want_baseband=True, want_power=want_power)
        return bb, pw


class MessageReceiver:
    """Batched p25 MessageReceiver: feed() takes [S][n] baseband samples and returns the events of all
    streams ordered by (stream, sample); per stream this is the order the reference's feed() yields."""

    def __init__(self, ctx: Context):
        self.ctx = ctx

    def feed(self, samples: np.ndarray) -> np.ndarray:
        self.ctx.decode(samples)
        return self.ctx.poll()

    def resync(self, stream: int):
        self.ctx.resync(stream)


class ReplayReceiver:
    """reference src/replay.rs:11-57 for one or more f32le/48 kHz/mono baseband recordings."""

    READ_BYTES = 32768  # src/replay.rs:27

    def __init__(self, n_streams: int = 1, device: int = 0, on_voice_frame=None):
        self.ctx = Context(n_streams, fmt=FMT_U8_IQ, decimation=5, max_chunk_samples=BUF_SAMPLES,
                           max_baseband=self.READ_BYTES // 4, device=device)
        self.msg = MessageReceiver(self.ctx)
        self.on_voice_frame = on_voice_frame
        self.events: list[np.ndarray] = []

    def replay(self, streams) -> np.ndarray:
        """streams: list of binary file objects (one per stream), read in 32,768-byte blocks (src/replay.rs:27).
        Every feed carries the same number of samples for every stream (the C ABI's shape), so a stream's unread
        remainder is kept for the next round; replay ends when the shortest recording ends.  Unlike src/replay.rs:36 a
        short final read is fed at its true length (the reference re-feeds the stale tail of its buffer)."""
        pend = [b"" for _ in streams]
        while True:
            pend = [p if len(p) >= self.READ_BYTES else p + f.read(self.READ_BYTES) for p, f in zip(pend, streams)]
            n = min(min(len(p) for p in pend), self.READ_BYTES) // 4
            if n == 0:
                break
            chunk = np.stack([np.frombuffer(p[: 4 * n], dtype="<f4") for p in pend])
            pend = [p[4 * n:] for p in pend]
            ev = self.msg.feed(chunk)
            if self.on_voice_frame is not None:
                for e in ev[ev["kind"] == EV_VOICE_FRAME]:
                    self.on_voice_frame(e)
            self.events.append(ev)
        return np.concatenate(self.events) if self.events else np.zeros(0, dtype=EVENT_DTYPE)
