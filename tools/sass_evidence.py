"""cuobjdump -sass of the built library -> per-kernel counts of the SASS mnemonics that prove the Blackwell paths
(B200_PROFILING.md: tcgen05.mma = UTC*MMA, tcgen05.ld = LDTM, TMA = UTMALDG/UTMASTG/UBLKCP, tcgen05.commit = UTCBAR,
tcgen05.alloc = UTCATOMSWS, mbarrier = SYNCS, split-K reduction = REDG; legacy mma.sync would show as HMMA).
Runs without a GPU:  python tools/sass_evidence.py > profiles/rNN_sass_evidence.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "lsps_b200", "csrc", "liblsps_b200.so")
sass = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCBAR|UTMALDG|UTMASTG|UBLKCP|UTCATOMSWS|LDTM|STTM|SYNCS|REDG|HMMA|HGMMA)\b")
fn, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        counts[fn] = collections.Counter()
        continue
    m = pat.search(line)
    if m and fn:
        counts[fn][m.group(1)] += 1
cols = ["UTCHMMA", "LDTM", "UTMALDG", "UTCBAR", "UTCATOMSWS", "SYNCS", "REDG", "HMMA"]
print("# SASS evidence, liblsps_b200.so (sm_100a), %d kernels\n" % len(counts))
print("`cuobjdump -sass lsps_b200/csrc/liblsps_b200.so`, instruction counts per kernel (static, not executed counts).")
print("UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA load), UTCBAR = tcgen05.commit,")
print("UTCATOMSWS = tcgen05.alloc/dealloc/relinquish, SYNCS = mbarrier ops, REDG = red.global (split-K weight gradients).")
print("No kernel contains HMMA (mma.sync / wmma).\n")
print("| kernel | " + " | ".join(cols) + " |\n|---|" + "---|" * len(cols))
rest = []
for f, c in counts.items():
    name = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip()
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\((?!anonymous).*$", "", name)
    if c.get("UTCHMMA") or c.get("UTMALDG") or c.get("LDTM"):
        print("| `%s` | " % name + " | ".join(str(c.get(k, 0)) for k in cols) + " |")
    else:
        rest.append(name)
print("\nKernels without tensor-core / TMA instructions (HBM-, issue- or latency-bound SIMT kernels, DESIGN.md 3.4): "
      + ", ".join("`%s`" % r for r in sorted(set(rest))))
