#!/usr/bin/env python3
"""sass_summary.py -- per-kernel SASS evidence of libp25cu.so (no GPU needed): instruction counts of the mnemonics that
tell a Blackwell-native streaming kernel from a recompiled one (B200_PROFILING.md: UBLKCP / UTMALDG = TMA bulk copies,
SYNCS = mbarrier, UCGABAR = cluster barrier, FFMA2 / FADD2 / FMUL2 = packed FP32, UTC*MMA / LDTM / STTM = tcgen05 + TMEM, HMMA / IMMA = legacy mma.sync, IMMA on the raw u8 IQ bytes), plus registers,
spills and shared memory from `cuobjdump -res-usage`.

    python tools/sass_summary.py [p25rx_b200/libp25cu.so] > profiles/sass_summary.txt
"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ["FFMA2", "FFMA", "FADD2", "FMUL2", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "HMMA", "IMMA",
         "UCGABAR_ARV", "UCGABAR_WAIT", "PRMT", "LDS", "STS", "ST", "LDG", "STG", "SHFL", "BAR", "REDUX", "ATOMG", "RED", "MUFU", "IDP"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "p25rx_b200", "libp25cu.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
    kernels, cur, arch = OrderedDict(), None, set()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = Counter()
            continue
        m = re.match(r"\s*arch = (\S+)", line)
        if m:
            arch.add(m.group(1))
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1).split(".")[0]
            kernels[cur]["_total"] += 1
            for w in WATCH:
                if op == w or (w == "RED" and op == "REDG"):
                    kernels[cur][w] += 1
    usage = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", res):
        usage[m.group(1)] = tuple(int(x) for x in m.groups()[1:])
    names = demangle(list(kernels))
    print(f"# {os.path.relpath(lib, ROOT)}  arch: {', '.join(sorted(arch))}  kernels: {len(kernels)}")
    tot = Counter()
    for k, c in kernels.items():
        tot.update(c)
    print("# totals: " + "  ".join(f"{w} {tot[w]}" for w in WATCH if tot[w]))
    print("# tcgen05 / TMEM (UTC*MMA, LDTM, STTM): " + str(sum(tot[w] for w in ("UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM"))) +
          "   legacy tensor (HMMA): " + str(tot["HMMA"]) + "   TMA bulk (UBLKCP + UTMALDG): " + str(tot["UBLKCP"] + tot["UTMALDG"]))
    print()
    for k, c in kernels.items():
        reg, stack, shared, local = usage.get(k, (0, 0, 0, 0))
        short = re.sub(r"\(.*", "", names.get(k, k))
        print(f"{short}\n    instr {c['_total']:5d}  regs {reg:3d}  stack {stack:4d}  static smem {shared:6d}  local {local}   " +
              "  ".join(f"{w} {c[w]}" for w in WATCH if c[w]))


if __name__ == "__main__":
    main()
