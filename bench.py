#!/usr/bin/env python3
"""bench.py -- headline benchmark of the P25 baseband hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): 1,024 independent synthetic P25 control-channel IQ streams per
GPU, 2.4 MS/s cf32 (configs[0]'s sample format), 150 ms = 360,000 samples per stream per step,
decimated by 50 to 48 kHz, C4FM-demodulated, frame-synchronised and TSBK-decoded.  One step is one
pass of the hot path (p25cu_process: ddc_fm kernel + decode walker) over that batch, 2.95 GB of input,
far larger than the 126 MB L2, so every step streams from HBM.  The signal is periodic and
phase-continuous, so consecutive steps are a continuous transmission and every step decodes real
TSDUs (checked: 8 events per stream per step, CRCs valid).

  value      whole-job IQ Msamples/s with the input already resident in HBM: device time (CUDA events on
             the library's stream) of K x p25cu_process plus one final event drain (p25cu_poll).
  e2e        the same metric through the public call sequence with HOST (pinned) input:
             each step = p25cu_process(host pointer) [H2D copy inside] + p25cu_poll [D2H of events].
  roofline   dominant kernel p25_ddc_fm_kernel: algorithmic bytes (8 + 4/50 per input sample) divided by
             its own CUDA-event time inside the timed region, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the oracle (C++ restatement of the reference chain, -march=native) on a bounded sample
             of the same workload using every host core.  `--impl reference` runs that arm alone.

Multi-GPU: streams shard by rank with no collective on the data path; rank r owns
streams [r*1024, (r+1)*1024).  torch.distributed (NCCL) is used only for the barrier and the MAX of the
per-rank device times.  scaling = weak.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "spec")):
    if p not in sys.path:
        sys.path.insert(0, p)

STREAMS_PER_GPU = 1024
FS = 2_400_000
DECIM = 50
N_PER_STEP = 360_000          # 150 ms = two 360-dibit TSDUs per stream per step
N_BASE = 16                   # distinct seeded transmissions; streams are circular shifts of them
METRIC = "Msamples/s IQ demod+decode (real-time P25 channels = baseband samples/s / 48000)"
ALG_BYTES_PER_SAMPLE = 8.0 + 4.0 / DECIM     # SURVEY.md section 8(d), DESIGN.md section 4


class Workload:
    """cfg2 (default, the configuration the metric is quoted on): 1,024 cf32 2.4 MS/s streams per GPU, weak scaling.
    cfg5 (BASELINE.json configs[4]): 65,536 streams in the reference's own format (u8 IQ at 240 kS/s, /5), sharded
    over the ranks (strong scaling), generated on the device from 16 seeded transmissions."""

    def __init__(self, name: str, world: int):
        self.name = name
        if name == "cfg5":
            assert 65536 % world == 0
            self.streams, self.fs, self.decim, self.fmt, self.n, self.scaling = 65536 // world, 240_000, 5, "u8", 36_000, "strong"
            self.bps = 2
            self.kernel = "fast5::p25_ddc5_fm_kernel<u8> (/5)"
            self.desc = (f"configs[4]: 65,536 synthetic P25 control-channel streams in the reference's own format (u8 IQ, "
                         f"240 kS/s, /5 -> 48 kHz), {self.streams} per GPU, 36000 samples (150 ms) per stream per step, "
                         "C4FM demod + frame sync + NID/TSBK decode")
        else:
            self.streams, self.fs, self.decim, self.fmt, self.n, self.scaling = STREAMS_PER_GPU, FS, DECIM, "cf32", N_PER_STEP, "weak"
            self.bps = 8
            self.kernel = "fast::p25_ddc_fm_stream_kernel (cf32, /50)"
            self.desc = ("configs[1]: 1024 synthetic P25 control-channel IQ streams per GPU, cf32 2.4 MS/s, 360000 samples "
                         "(150 ms) per stream per step, /50 -> 48 kHz, C4FM demod + frame sync + NID/TSBK decode")
        self.alg_bytes_per_sample = self.bps + 4.0 / self.decim

    def config(self, world: int):
        gb = self.streams * self.n * self.bps / 1e9
        return {"workload": self.desc, "streams_per_gpu": self.streams, "samples_per_stream_per_step": self.n,
                "sample_rate": self.fs, "decimation": self.decim, "input_format": self.fmt, "snr_db": 20,
                "l2_policy": f"inputs ({gb:.2f} GB/step) larger than L2",
                "parallelism": f"streams sharded over {world} GPU(s), no collective"}

    def base(self):
        from tools import p25tx as tx
        rows = []
        for b in range(N_BASE):
            st = tx.control_channel(1000 + b, 2, lead_idle=0)
            assert len(st.dibits) * 10 * self.decim == self.n
            iq = tx.modulate_iq_periodic(st.dibits, self.fs, snr_db=20.0, cfo_cycles=3 * (b - N_BASE // 2), seed=b)
            rows.append(tx.iq_to_u8(iq).reshape(self.n, 2) if self.fmt == "u8" else iq.view(np.float32).reshape(self.n, 2))
        return np.stack(rows)

    def oracle_input(self, n_streams: int):
        """Host array of the first n_streams streams in the oracle's layout."""
        base = self.base()
        out = np.empty((n_streams,) + base.shape[1:], dtype=base.dtype)
        for s in range(n_streams):
            out[s] = np.roll(base[s % N_BASE], -((s // N_BASE) * 5003 % self.n), axis=0)
        return out


def make_base_streams():
    from tools import p25tx as tx
    base = []
    for b in range(N_BASE):
        st = tx.control_channel(1000 + b, 2, lead_idle=0)
        assert len(st.dibits) * 10 * DECIM == N_PER_STEP
        base.append(tx.modulate_iq_periodic(st.dibits, FS, snr_db=20.0, cfo_cycles=3 * (b - N_BASE // 2), seed=b))
    return np.stack(base)


def fill_streams(dst: np.ndarray, base: np.ndarray, first_stream: int):
    """dst[s] = circular shift of base[(first_stream + s) % N_BASE]; shift depends on the global stream id."""
    for s in range(dst.shape[0]):
        g = first_stream + s
        sh = (g // N_BASE) * 5003 % N_PER_STEP
        src = base[g % N_BASE]
        dst[s, : N_PER_STEP - sh] = src[sh:]
        dst[s, N_PER_STEP - sh:] = src[:sh]


class ClockSampler(threading.Thread):
    """NVML poll of SM clock / throttle reasons during the timed region."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def result(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class StreamGate:
    """Holds a CUDA stream back until the host opens it (cuStreamWaitValue32 on a pinned flag).  The timed loop is
    enqueued behind the gate, so the device runs its K steps back to back from a full launch queue -- what a
    steady-state producer sees -- instead of inheriting the scheduling jitter of eight Python ranks sharing one host.
    Falls back to no gate if the driver API is unavailable."""

    def __init__(self, stream_handle: int, enabled: bool = True):
        self.ok = False
        # A closed gate deadlocks anything that makes launches synchronous (a profiler replaying kernels one by one,
        # CUDA_LAUNCH_BLOCKING): never gate then, and a watchdog opens the gate after 250 ms whatever happens.
        blocking = os.environ.get("CUDA_LAUNCH_BLOCKING", "0") not in ("", "0")
        injected = any(k.startswith(("CUDA_INJECTION", "NV_COMPUTE_PROFILER", "NSIGHT", "NV_NSIGHT", "NVTX_INJECTION")) for k in os.environ)
        if not enabled or blocking or injected:
            return
        try:
            import torch
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                from cuda.bindings import driver as drv
            self.flag = torch.zeros(1, dtype=torch.int32).pin_memory()
            err, = drv.cuStreamWaitValue32(drv.CUstream(stream_handle), drv.CUdeviceptr(self.flag.data_ptr()), 1,
                                           drv.CUstreamWaitValue_flags.CU_STREAM_WAIT_VALUE_GEQ)
            self.ok = int(err) == 0
            if self.ok:
                self.watchdog = threading.Timer(0.25, self.open)
                self.watchdog.daemon = True
                self.watchdog.start()
        except Exception:
            self.ok = False

    def open(self):
        if self.ok:
            self.flag[0] = 1


def pin_to_gpu_numa_node(index: int) -> None:
    """Multi-GPU host leg: run this rank on the CPUs next to its GPU so that the pinned staging buffer is allocated
    (first touch) on that NUMA node and eight ranks do not all stream from one socket's memory.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def single_stream_split(po, native: bool, fmt: int, front: bool, row: np.ndarray, decim: int) -> dict:
    """The reference's actual shape: one stream on one thread, timed separately for the `demod` thread's
    work (src/demod.rs:62-119) and the `receiver` thread's (src/recv.rs:140-167), a few repetitions of one step's row."""
    reps = 4
    chain, rx = po.DemodChain(fmt, front, native=native), po.MessageReceiver(native=native)
    t0 = time.perf_counter()
    bbs = [chain.feed(row) for _ in range(reps)]
    t1 = time.perf_counter()
    for bb in bbs:
        rx.feed(bb)
    t2 = time.perf_counter()
    n_in = reps * (row.size // (2 if fmt == po.FMT_U8 else 1))
    return {"threads": 1, "demod_iq_msamples_per_s": n_in / (t1 - t0) / 1e6,
            "decode_baseband_msamples_per_s": n_in / decim / (t2 - t1) / 1e6,
            "realtime_channels_one_core": 1.0 / ((t2 - t0) / (n_in / decim / 48000.0))}


def cpu_arm(steps: int, warmup: int, iq: np.ndarray | None = None, wl: "Workload | None" = None):
    """Oracle CPU implementation on the host cores (rank 0 only).  Returns (Msamples/s, ms/step, info).
    Each step is the full 1,024-stream batch of the GPU arm (about 8 core-seconds of work); for cfg5 a bounded
    sample of 4,096 of the 65,536 streams."""
    if wl is not None and wl.name == "cfg5":
        return cpu_arm_cfg5(steps, warmup, wl)
    from oracle import pyoracle as po
    try:
        po.build(native=True)
        native = True
    except Exception:
        native = False
    cores = os.cpu_count() or 1
    L = po.lib(native)
    L.p25o_set_always_correlate(1)       # the reference correlates on every sample [RECALL], see oracle header
    n_streams = STREAMS_PER_GPU
    if iq is None:
        iq = np.empty((n_streams, N_PER_STEP), dtype=np.complex64)
        fill_streams(iq, make_base_streams(), 0)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        total, counts, _ = po.batch_run(po.FMT_CF32, True, iq, n_streams, N_PER_STEP, cores, native=native)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        assert total >= 3 * n_streams, "oracle decoded too few events"   # fresh receivers: >= 1 whole TSDU each
    ms = 1e3 * float(np.mean(times))
    val = n_streams * N_PER_STEP / (ms * 1e-3) / 1e6
    info = {"value": val, "unit": "Msamples/s", "cores": cores, "kind": "port",
            "sample": f"all {n_streams} streams x {N_PER_STEP} samples per step ({steps} timed step(s)), {cores} threads, "
                      f"oracle built {'-march=native' if native else '-march=x86-64-v3'}; the Rust reference cannot be built here",
            "single_stream": single_stream_split(po, native, po.FMT_CF32, True, iq[0], DECIM)}
    return val, ms, info


def cpu_arm_cfg5(steps: int, warmup: int, wl: "Workload"):
    from oracle import pyoracle as po
    try:
        po.build(native=True)
        native = True
    except Exception:
        native = False
    cores = os.cpu_count() or 1
    po.lib(native).p25o_set_always_correlate(1)
    n_streams = 4096
    iq = wl.oracle_input(n_streams)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        total, _, _ = po.batch_run(po.FMT_U8, False, iq, n_streams, wl.n, cores, native=native)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        assert total >= 3 * n_streams
    ms = 1e3 * float(np.mean(times))
    val = n_streams * wl.n / (ms * 1e-3) / 1e6
    return val, ms, {"value": val, "unit": "Msamples/s", "cores": cores, "kind": "port",
                     "sample": f"{n_streams} of the 65536 streams x {wl.n} u8 samples per step ({steps} timed step(s)), {cores} "
                               f"threads, oracle built {'-march=native' if native else '-march=x86-64-v3'}; the Rust reference cannot be built here",
                     "single_stream": single_stream_split(po, native, po.FMT_U8, False, iq[0], wl.decim)}


def config_dict(n_gpus: int):
    return {"workload": "configs[1]: 1024 synthetic P25 control-channel IQ streams per GPU, cf32 2.4 MS/s, 360000 samples "
                        "(150 ms) per stream per step, /50 -> 48 kHz, C4FM demod + frame sync + NID/TSBK decode",
            "streams_per_gpu": STREAMS_PER_GPU, "samples_per_stream_per_step": N_PER_STEP, "sample_rate": FS,
            "decimation": DECIM, "input_format": "cf32", "snr_db": 20, "l2_policy": "inputs (2.95 GB/step) larger than L2",
            "parallelism": f"streams sharded over {n_gpus} GPU(s), no collective"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = Workload(args.workload, max(1, args.gpus))
    val, ms, info = cpu_arm(args.steps, args.warmup, wl=wl)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": wl.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": wl.config(args.gpus), "cpu_baseline": info,
            "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "realtime_channels": val * 1e6 / wl.decim / 48000.0, "gpu_launches": 0}
    print(json.dumps(line))


def run_b200(args):
    import torch
    import p25rx_b200 as p25

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    full_affinity = os.sched_getaffinity(0)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        pin_to_gpu_numa_node(local)          # pinned staging memory on the GPU's own NUMA node (first touch)

    wl = Workload(args.workload, world)
    S, n, K, W = wl.streams, wl.n, args.steps, max(args.warmup, 3)
    # streams = circular shifts of N_BASE seeded transmissions, tiled on the device, then
    # mirrored into pinned host memory for the end-to-end leg
    from tools.shape_bench import tile_on_device
    dev = tile_on_device(torch.from_numpy(wl.base()).cuda(), S, first=rank * S)
    host = torch.empty(dev.shape, dtype=dev.dtype, pin_memory=True)
    host.copy_(dev)
    torch.cuda.synchronize()          # the library runs on its own stream: inputs complete before they are handed over
    slots = 8 * (W + K) + 32
    ctx = p25.Context(S, fmt=p25.FMT_U8_IQ if wl.fmt == "u8" else p25.FMT_CF32_IQ, decimation=wl.decim, max_chunk_samples=n,
                      device=local, event_slots=slots)
    stream = torch.cuda.ExternalStream(ctx.cuda_stream, device=local)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- leg 1: device-resident input ("value") + per-kernel timing for the roofline
    for _ in range(W):
        ctx.process(dev, n)
    ctx.sync()
    ctx.poll(copy=False)
    sampler = ClockSampler(local)
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ctx.launch_count
    ctx.demod_timing(True)
    barrier()
    sampler.start()
    t_host0 = time.perf_counter()
    with torch.cuda.stream(stream):
        gate = StreamGate(ctx.cuda_stream, enabled=not args.no_gate)   # the device starts when the launch queue holds the first steps
        t_begin.record(stream)
        for k in range(K):
            if k == 48:
                gate.open()                              # bounded: never let a gated queue fill up
            ctx.process(dev, n)                          # demod kernel, then the decode walker (beside the next demod kernel)
        gate.open()
        enqueue_ms = 1e3 * (time.perf_counter() - t_host0)
        events = ctx.poll(copy=False)                    # event compaction + D2H into pinned memory (synchronises)
        t_end.record(stream)
    barrier()
    sampler.stop_flag = True
    sampler.join()
    launches = ctx.launch_count - l0
    dev_ms = t_begin.elapsed_time(t_end)
    ddc_ms, n_timed = ctx.demod_timing(False)                       # the library's own event pair right around each kernel launch
    assert n_timed == min(K, 128)
    events = events.copy()
    # per-kernel breakdown with the two kernels serialised on one stream (not part of `value`)
    ctx.set_overlap(False)
    bk = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(5)]
    with torch.cuda.stream(stream):
        for e3 in bk:
            e3[0].record(stream)
            ctx.demod(dev, n, want_baseband=False)
            e3[1].record(stream)
            ctx.decode()
            e3[2].record(stream)
    ctx.sync()
    ddc_serial_ms = sum(e[0].elapsed_time(e[1]) for e in bk) / len(bk)
    walk_ms = sum(e[1].elapsed_time(e[2]) for e in bk) / len(bk)
    ctx.poll(copy=False)
    ctx.set_overlap(wl.decim == 50)                  # back to the library's default for this shape
    # correctness of the timed work: 2 TSDUs (2 NIDs + 6 TSBKs) per stream per step, all CRCs valid
    n_tsbk = int(np.count_nonzero(events["kind"] == p25.EV_TSBK))
    n_err = int(np.count_nonzero(events["kind"] == p25.EV_ERROR))
    assert n_tsbk >= 6 * S * K - 6 * S and n_err <= S, (n_tsbk, n_err, len(events))
    import p25_spec as SP
    for e in events[events["kind"] == p25.EV_TSBK][:: max(1, n_tsbk // 2000)]:
        pl = bytes(e["payload"][:12])
        assert SP.crc_ccitt_p25(pl[:10]) == (pl[10] << 8 | pl[11]), "decoded TSBK fails its CRC"

    # ---------------- leg 2: end to end through the public calls with pinned host input
    host_np = host.numpy()
    for _ in range(2):
        ctx.process(host_np, n)
        ctx.poll(copy=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d2h = 0
    barrier()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for k in range(K):
            ctx.process(host_np, n)                      # H2D copy of the step's input happens inside
            got = ctx.poll(copy=False)                   # D2H of the step's events
            d2h += got.nbytes + 8
        e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1)

    t = torch.tensor([dev_ms, e2e_ms, ddc_ms, walk_ms], dtype=torch.float64, device="cuda")  # MAX over ranks
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, ddc_ms, walk_ms = [float(x) for x in t.tolist()]
    total_samples = world * S * n * K
    value = total_samples / (dev_ms * 1e-3) / 1e6
    e2e_value = total_samples / (e2e_ms * 1e-3) / 1e6

    peak, peak_src = measured_peak_gbs()
    alg_bytes = S * n * wl.alg_bytes_per_sample
    achieved = alg_bytes / (ddc_ms * 1e-3) / 1e9
    line = {"metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": wl.scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": wl.config(world),
            "realtime_channels": value * 1e6 / wl.decim / 48000.0,
            "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": S * n * wl.bps, "d2h_bytes_per_step": d2h // K,
                    "ms_per_step": e2e_ms / K},
            "gpu_launches": int(launches),
            "kernels": {"p25_ddc_fm_kernel_ms": ddc_ms, "step_ms": dev_ms / K, "host_enqueue_ms_per_step": enqueue_ms / K,
                        "serialised": {"p25_ddc_fm_kernel_ms": ddc_serial_ms, "p25_walk_kernel_ms": walk_ms}},
            "roofline": {"kernel": wl.kernel, "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes},
            "clocks": sampler.result(), "events_checked": {"tsbk": n_tsbk, "errors": n_err}}
    try:
        with open(os.path.join(ROOT, "profiles", "ddc_fm_traffic.json" if wl.name != "cfg5" else "ddc5_u8_traffic.json")) as f:
            tr = json.load(f)
            line["roofline"]["traffic"] = tr.get("dram_bytes_per_launch")
    except Exception:
        pass
    ctx.close()
    if rank == 0:
        os.sched_setaffinity(0, full_affinity)
        if world == 1 and not args.no_cpu:
            if wl.name == "cfg5":
                _, _, info = cpu_arm(1, 0, wl=wl)
            else:
                _, _, info = cpu_arm(1, 0, host.numpy().view(np.complex64).reshape(S, n))
            line["cpu_baseline"] = info
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-gate", action="store_true", help="do not hold the stream back while the timed steps are enqueued")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg5"],
                    help="cfg2 (default): 1,024 cf32 2.4 MS/s streams per GPU; cfg5: 65,536 u8 240 kS/s streams over all GPUs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
