"""nvcc -Xptxas -v over lsps_b200/csrc/*.cu -> registers / spills / static shared memory per kernel (no GPU needed).
python tools/ptxas_table.py > profiles/rNN_ptxas_resources.md"""
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = []
for src in sorted(glob.glob(os.path.join(ROOT, "lsps_b200", "csrc", "*.cu"))):
    out = subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17",
                          "-Xptxas", "-v", "-c", src, "-o", "/dev/null"], capture_output=True, text=True).stderr
    fn = None
    spill = "0/0"
    for line in out.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r"\(anonymous namespace\)::", "", fn)
            fn = re.sub(r"^void ", "", fn)
            fn = re.sub(r"\((?!anonymous).*$", "", fn)
            continue
        m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            spill = "%s/%s" % (m.group(1), m.group(2))
        m = re.search(r"Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?", line)
        if m and fn:
            rows.append((os.path.basename(src), fn, m.group(1), spill, m.group(3) or "0"))
            fn = None
print("# ptxas resource usage, sm_100a (static shared memory only; the tcgen05 kernels add 190-200 KB dynamic)\n")
print("| file | kernel | registers | spill st/ld bytes | static smem |\n|---|---|---|---|---|")
for r in rows:
    print("| %s | `%s` | %s | %s | %s |" % r)
