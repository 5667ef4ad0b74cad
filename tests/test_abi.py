"""The C-ABI library loads on a CPU-only machine and exports every symbol include/p25cu.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "p25cu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(p25cu_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from p25rx_b200 import _lib
    _lib.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), n
    assert sorted(_lib.EXPORTS) == names


def test_struct_layouts_match_header():
    from p25rx_b200 import _lib
    assert ctypes.sizeof(_lib.Config) == 40
    assert _lib.EVENT_DTYPE.itemsize == 80 and _lib.EVENT_DTYPE.fields["payload"][1] == 20
    assert ctypes.sizeof(_lib.Stats) == 12 * 4 * 8


def test_create_fails_loudly_without_gpu():
    """No CPU fallback: without a usable sm_100 device create returns an error code and a message."""
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    import p25rx_b200 as p
    with pytest.raises(p.P25Error) as ei:
        p.Context(4)
    assert ei.value.status == -2


def test_bad_config_is_rejected():
    from p25rx_b200 import _lib
    L = _lib.lib()
    h = ctypes.c_void_p()
    cfg = _lib.Config(0, 4, 1, 7, 1000, 0, _lib.ABI_VERSION, 0)     # decimation 7 is not supported
    assert L.p25cu_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    assert b"bad config" in L.p25cu_last_error(None)


def test_product_never_touches_the_oracle():
    """Nothing under p25rx_b200/ or include/ may include, import, link or load oracle/ (tests and bench only)."""
    bad = re.compile(r"(#\s*include[^\n]*oracle|^\s*(from|import)\s+oracle|libp25oracle|pyoracle|-lp25oracle)", re.M)
    for base in ("p25rx_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                    assert not bad.search(open(os.path.join(dp, f), errors="ignore").read()), os.path.join(dp, f)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (CPU oracle, no GPU needed): one JSON line with the keys the driver reads."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["vs_baseline"] is None and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "single_stream" in d["cpu_baseline"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
