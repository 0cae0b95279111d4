#!/usr/bin/env python
"""Per-kernel SASS evidence for libpodb200.so (runs without a GPU): counts of the instructions that show which
hardware path each kernel uses.

  DMMA     fp64 tensor-core MMA (mma.sync.m8n8k4.f64 -> DMMA.8x8x4; tcgen05 has no fp64 kind)
  UTMALDG  TMA tensor load (cp.async.bulk.tensor)        UBLKCP  TMA bulk copy (cp.async.bulk)
  SYNCS    mbarrier operations                            LDGSTS  cp.async
  DFMA     fp64 CUDA-core FMA                             MUFU    special-function unit

usage: python tools/sass_summary.py [libpodb200.so] > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "pod_uqnn_b200", "libpodb200.so")
elfs = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
archs = collections.Counter(re.findall(r"\.(sm_\w+)\.cubin", elfs))
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEYS = ["DMMA", "UTMALDG", "UBLKCP", "SYNCS", "LDGSTS", "DFMA", "MUFU"]
rows, name, cnt = [], None, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        if name:
            rows.append((name, cnt))
        name, cnt = m.group(1), collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        cnt["instr"] += 1
        op = m.group(1).split(".")[0]
        if op in KEYS:
            cnt[op] += 1
if name:
    rows.append((name, cnt))
names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()


def short(n):
    n = re.sub(r"\(anonymous namespace\)::", "", n)
    n = re.sub(r"^void ", "", n)
    m = re.match(r"([\w:]+(?:<[^()]*>)?)\(", n)
    return m.group(1) if m else n


print(f"# {os.path.basename(so)}: {sum(archs.values())} cubins, " + ", ".join(f"{v} x {k}" for k, v in archs.items()))
print(f"# {len(rows)} kernels; totals: " + ", ".join(f"{k} {sum(c[k] for _, c in rows)}" for k in KEYS))
print(f"{'kernel':<78}" + "".join(f"{k:>8}" for k in KEYS) + f"{'instr':>8}")
for n, (_, c) in sorted(zip(map(short, names), rows)):
    print(f"{n[:77]:<78}" + "".join(f"{c[k]:>8}" for k in KEYS) + f"{c['instr']:>8}")
