"""Reads an `ncu --set full` report here (no GPU needed) and prints one markdown row per profiled launch:
   python tools/ncu_summary.py gpurun_out/prof_targets.ncu-rep [labels.md] > profiles/rNN_ncu_<what>.md
labels.md: optional stdout of tools/ncu_targets.py (its row order = launch order of the single-kernel cases)."""
import csv
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rd MB"), ("dram__bytes_write.sum", "wr MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("lts__t_bytes.sum", "L2 MB"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor % active"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor % elapsed"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    ki = idx["Kernel Name"]
    print("| # | kernel | " + " | ".join(c[1] for c in COLS) + " | dram GB/s |")
    print("|---|---|" + "---|" * (len(COLS) + 1))
    for n, r in enumerate(data):
        vals = []
        d = {}
        for key, _ in COLS:
            i = idx.get(key)
            v = r[i] if i is not None else ""
            u = units[i] if i is not None else ""
            try:
                f = float(v.replace(",", ""))
                if u == "byte":
                    f /= 1e6
                elif u == "Kbyte":
                    f /= 1e3
                elif u == "Gbyte":
                    f *= 1e3
                elif u in ("ns", "nsecond"):
                    f /= 1e3
                elif u in ("ms", "msecond"):
                    f *= 1e3
                d[key] = f
                vals.append("%.1f" % f if f < 1e5 else "%.0f" % f)
            except ValueError:
                vals.append(v)
        name = r[ki].split("(")[0].replace("void ", "")
        gbs = ""
        try:
            gbs = "%.0f" % ((d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]) / d["gpu__time_duration.sum"] * 1e3)
        except Exception:  # noqa
            pass
        print("| %d | %s | %s | %s |" % (n, name, " | ".join(vals), gbs))


if __name__ == "__main__":
    main()
