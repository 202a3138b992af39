#!/usr/bin/env python
"""Per-kernel SASS digest of libb200gym.so: code size, and the mnemonics that prove which hardware paths a kernel uses
(UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st (tensor memory), UBLKCP = cp.async.bulk (TMA engine), UTCBAR = tcgen05.commit,
SYNCS = mbarrier, SHFL = warp shuffles, REDUX = __reduce_*_sync, LDL / STL = local memory, BAR = block barriers).
Usage: python tools/sass_digest.py [lib] > profiles/r02_sass_digest.txt      (needs cuobjdump; no GPU)"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "isaacgymenv_b200/lib/libb200gym.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "FFMA", "DFMA", "MUFU", "SHFL", "REDUX", "LDL", "STL", "LDS", "STS", "LDG", "STG", "ATOM", "RED", "BAR"]
fn, counts, sizes = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = m.group(1)
        counts[fn] = collections.Counter()
        sizes[fn] = 0
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if fn and m:
        sizes[fn] = max(sizes[fn], int(m.group(1), 16) + 16)
        op = m.group(3).split(".")[0]
        counts[fn]["_n"] += 1
        for k in KEYS:
            if op == k or (k in ("ATOM", "RED") and op in (k, k + "G", k + "S")):
                counts[fn][k] += 1
demangle = subprocess.run(["c++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# SASS digest of {lib} (sm_100a), cuobjdump -sass; columns: instructions, code bytes, then mnemonic counts (zeros omitted)")
for (fn, c), name in sorted(zip(counts.items(), demangle), key=lambda t: -t[0][1]["_n"]):
    name = re.sub(r"\(anonymous namespace\)::|b2g::", "", name)
    name = re.sub(r"\(.*", "", name)
    tags = " ".join(f"{k}={c[k]}" for k in KEYS if c[k])
    print(f"{name:<58s} {c['_n']:7d} instr {sizes[fn]:8d} B  {tags}")
