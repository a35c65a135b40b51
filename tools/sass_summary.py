#!/usr/bin/env python
"""Per-kernel SASS evidence: ``cuobjdump -sass cppf_b200/libcppf_b200.so`` -> ``profiles/sass_summary.md``.

Counts, per kernel, the mnemonics that prove which hardware paths the build uses (B200_PROFILING.md):
  UTCHMMA / UTCQMMA ... tcgen05.mma          LDTM / STTM ... tcgen05.ld / st (TMEM)       UTCBAR ... tcgen05.commit
  UTCCP ... tcgen05.cp                       UTMALDG / UTMASTG ... TMA tensor-map loads / stores (cp.async.bulk.tensor)
  UBLKCP ... cp.async.bulk (1-D TMA)         SYNCS ... mbarrier ops
  ATOMS ... shared-memory atomics            ATOMG / REDG / RED ... global atomics / reductions
Run by ``__graft_entry__.build()`` after every rebuild, so the committed file always describes the committed sources.
"""
from __future__ import annotations

import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cppf_b200", "libcppf_b200.so")
OUT = os.path.join(ROOT, "profiles", "sass_summary.md")
COLS = ["UTCHMMA", "UTCQMMA", "UTCCP", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "ATOMS", "ATOMG",
        "REDG", "RED", "LDS", "STS", "LDG", "STG", "MUFU", "FFMA", "DFMA"]


def demangle(names):
    try:
        r = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, timeout=60)
        out = r.stdout.strip().splitlines()
        if len(out) == len(names):
            return out
    except Exception:
        pass
    return names


def summarize(lib=LIB):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    ins = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)")
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = ins.match(line)
        if m:
            cur[m.group(1)] += 1
            cur["_total"] += 1
    return kernels


def short(name, width=110):
    name = re.sub(r"\s+", " ", name)
    return name if len(name) <= width else name[:width - 3] + "..."


def write(out=OUT, lib=LIB):
    k = summarize(lib)
    names = demangle(list(k.keys()))
    cols = [c for c in COLS if any(v[c] for v in k.values())]
    tot = collections.Counter()
    lines = ["# SASS summary of `cppf_b200/libcppf_b200.so` (sm_100a)", "",
             "Written by `tools/sass_summary.py` (run by `__graft_entry__.build()`): static instruction counts per kernel from",
             "`cuobjdump -sass`.  UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTMALDG / UTMASTG = TMA",
             "tensor-map load / store, UBLKCP = cp.async.bulk, SYNCS = mbarrier, ATOMS = shared-memory atomic, REDG / ATOMG =",
             "global reduction / atomic.", "",
             "| kernel | instr | " + " | ".join(cols) + " |", "|---|---|" + "---|" * len(cols)]
    for (mangled, c), nm in sorted(zip(k.items(), names), key=lambda t: -t[0][1]["_total"]):
        if not any(c[x] for x in cols if x not in ("LDG", "STG", "FFMA", "MUFU", "LDS", "STS")) and c["_total"] < 200:
            continue
        lines.append(f"| `{short(nm)}` | {c['_total']} | " + " | ".join(str(c[x]) if c[x] else "" for x in cols) + " |")
        tot.update(c)
    for c in k.values():
        pass
    all_tot = collections.Counter()
    for c in k.values():
        all_tot.update(c)
    lines += ["", "Library totals (all kernels): " + ", ".join(f"{x} {all_tot[x]}" for x in cols if all_tot[x]) +
              f"; {len(k)} kernels, {all_tot['_total']} instructions."]
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    return out, all_tot


if __name__ == "__main__":
    path, tot = write()
    print(path, {x: tot[x] for x in COLS if tot[x]}, file=sys.stderr)
