"""Turns ncu outputs into the markdown summaries kept under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches_TAG.csv            > profiles/TAG_launches.md
    python tools/summarize_ncu.py full gpurun_out/full_TAG.ncu-rep                > profiles/TAG_full.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

UNIT_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    d = collections.defaultdict(list)
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        rec = dict(zip(hdr, r))
        if rec["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(rec["Metric Value"].replace(",", "")) * UNIT_US[rec["Metric Unit"]]
        d[re.sub(r"\(.*", "", rec["Kernel Name"])].append(v)
    tot = sum(sum(v) for v in d.values())
    print("| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print(f"| `{k}` | {len(v)} | {sum(v) / 1e3:.3f} | {sum(v) / len(v):.1f} | {sum(v) / tot:.3f} |")
    print(f"\nTotal {tot / 1e3:.2f} ms over {sum(len(v) for v in d.values())} launches "
          "(per-launch times are cold-cache and serialised: compare shares).")


FULL = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps act %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("smsp__inst_executed.sum", "warp insts"),
]


def full(path):
    if path.endswith(".csv"):  # already the raw page
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print("| kernel | " + " | ".join(n for _, n in FULL) + " |\n|---|" + "---|" * len(FULL))
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")])
        cells = []
        for m, _ in FULL:
            if m in hdr:
                i = hdr.index(m)
                try:
                    cells.append(f"{float(r[i]):.4g} {units[i]}".strip())
                except ValueError:
                    cells.append(r[i])
            else:
                cells.append("-")
        print(f"| `{name}` | " + " | ".join(cells) + " |")
    print("\n(ncu --set full --clock-control none; caches flushed before every replay pass, so L2 carry-over "
          "between kernels is not visible here.)")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
