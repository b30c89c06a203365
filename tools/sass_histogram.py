#!/usr/bin/env python
"""SASS evidence for profiles/: per-kernel counts of the Blackwell-native mnemonics in libmvuld_b200.so.

    python tools/sass_histogram.py > profiles/r2_sass_histogram.md

UTCHMMA = tcgen05.mma (UTCQMMA etc. would be other kinds), UTMALDG / UTMASTG = TMA tensor load / store, LDTM / STTM =
tcgen05.ld / st (TMEM), UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA = legacy mma.sync (must be absent).
"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mvuld_b200", "libmvuld_b200.so")
MNEMONICS = ["UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "MUFU", "FFMA2", "UCGABAR_ARV",
             "MAPA"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        per[cur]["_total"] += 1
        for mn in MNEMONICS:
            if op == mn or op.startswith(mn + ".") or op.startswith(mn + "_"):
                per[cur][mn] += 1
def demangle(n):
    r = subprocess.run(["cu++filt", n], capture_output=True, text=True)
    s = r.stdout.strip() or n
    s = s.replace("(int)", "").replace("(bool)", "")
    s = re.sub(r"\(.*", "", s).replace("void ", "").replace("mv::", "")
    return s[:90]
tot = collections.Counter()
for c in per.values():
    tot.update(c)
print("# SASS mnemonic histogram of `mvuld_b200/libmvuld_b200.so` (sm_100a)\n")
print("`python tools/sass_histogram.py` (cuobjdump -sass).  UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store,")
print("LDTM / STTM = tcgen05.ld / st (TMEM), UTCBAR = tcgen05.commit, SYNCS = mbarrier, UCGABAR_ARV = barrier.cluster.arrive, HMMA = legacy")
print("mma.sync (absent).\n")
print(f"{len(per)} kernels, {tot['_total']} SASS instructions.  Library totals: " +
      ", ".join(f"{m} {tot[m]}" for m in MNEMONICS) + "\n")
print("| kernel | instr | " + " | ".join(MNEMONICS) + " |")
print("|---|---:|" + "---:|" * len(MNEMONICS))
rows = [(n, c) for n, c in per.items() if any(c[m] for m in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UCGABAR_ARV"))]
for n, c in sorted(rows, key=lambda kv: -kv[1]["UTCHMMA"]):
    print(f"| `{demangle(n)}` | {c['_total']} | " + " | ".join(str(c[m]) if c[m] else "" for m in MNEMONICS) + " |")
others = len(per) - len(rows)
print(f"\n{others} further kernels (row / graph / reduction kernels on CUDA cores) hold none of the tensor / TMA / TMEM mnemonics.")
