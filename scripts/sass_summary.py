"""Static SASS instruction classes per kernel of libknpemi_b200.so -> profiles/<tag>_sass_summary.md
Usage: python scripts/sass_summary.py <tag>   (needs cuobjdump and c++filt on PATH; no GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
so = os.path.join(ROOT, "knp-emi-cgx_b200", "libknpemi_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
CLASSES = [("UBLKCP (TMA bulk copy)", r"\bUBLKCP"), ("SYNCS (mbarrier)", r"\bSYNCS"), ("LDGSTS (cp.async)", r"\bLDGSTS"),
           ("SHFL", r"\bSHFL"), ("DFMA+DADD+DMUL", r"\b(DFMA|DADD|DMUL)\b"), ("MUFU", r"\bMUFU"),
           ("LDL+STL (local memory)", r"\b(LDL|STL)\b"), ("BAR", r"\bBAR\b"), ("ATOM/RED", r"\b(ATOM|ATOMG|RED|ATOMS)\b"),
           ("tensor-core ops", r"\b(HMMA|IMMA|DMMA|QMMA|UTCHMMA|UTCQMMA|UTCIMMA|TCGEN|UTCMMA)")]
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur and re.match(r"\s*/\*[0-9a-f]{4}\*/", line):
        counts[cur]["n"] += 1
        for name, rx in CLASSES:
            if re.search(rx, line):
                counts[cur][name] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()


def short(n):
    n = re.sub(r"^void ", "", n).replace("knp::", "").replace("(anonymous namespace)::", "")
    m = re.match(r"([A-Za-z_0-9]+)(<[^(]*>)?", n)
    return (m.group(1) + (m.group(2) or "")) if m else n


rows = sorted((short(n), c) for n, c in zip(names, counts.values()))
out = [f"# SASS evidence ({tag}): instruction classes per kernel of libknpemi_b200.so", "",
       "`python scripts/sass_summary.py` = `cuobjdump -sass knp-emi-cgx_b200/libknpemi_b200.so`, counted per function (static "
       "instruction counts).  UBLKCP = 1-D TMA bulk copy (`cp.async.bulk`), SYNCS = mbarrier arrive / try_wait, LDGSTS = `cp.async` "
       "(global to shared without registers), BAR = CTA barrier, ATOM/RED = atomics (setup kernels of amg_device.cu and the "
       "peer-memory flag handshakes only: integer counters / maxima, never a floating-point sum).  No tensor-core instruction "
       "anywhere: the path is fp64, sparse and bandwidth- or issue-bound.", "",
       "| kernel | SASS instructions | " + " | ".join(n for n, _ in CLASSES) + " |", "|---|---|" + "---|" * len(CLASSES)]
for n, c in rows:
    out.append(f"| {n} | {c['n']} | " + " | ".join(str(c[k]) for k, _ in CLASSES) + " |")
tot = collections.Counter()
for _, c in rows:
    tot.update(c)
out.append(f"| **total ({len(rows)} kernels)** | {tot['n']} | " + " | ".join(str(tot[k]) for k, _ in CLASSES) + " |")
path = os.path.join(ROOT, "profiles", f"{tag}_sass_summary.md")
open(path, "w").write("\n".join(out) + "\n")
print(path, len(rows), "kernels", dict((k, tot[k]) for k, _ in CLASSES))
