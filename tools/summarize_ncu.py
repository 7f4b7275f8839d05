"""Turns ncu outputs brought back in gpurun_out/ into the small markdown summaries committed under profiles/.

    python tools/summarize_ncu.py launches <launches.csv> [--from-last <kernel substring>]   # per-kernel time + share
    python tools/summarize_ncu.py full <file.ncu-rep>                                        # key metrics per captured launch
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "IPC per SM"),
]


def launches(path, from_last=None):
    lines = [l for l in open(path) if not l.startswith("==")]
    seq = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        seq.append((row["Kernel Name"].split("(")[0].replace("void ", ""), v, row.get("Grid Size", ""), row.get("Block Size", "")))
    if from_last:
        idx = [i for i, s in enumerate(seq) if from_last in s[0]]
        seq = seq[idx[-1]:]
    agg = collections.OrderedDict()
    for n, v, g, b in seq:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v for _, v, _, _ in seq)
    print(f"| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {c} | {v:.1f} | {100 * v / tot:.1f} % |")
    print(f"| **all** | {len(seq)} | {tot:.1f} | 100 % |")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("| launch | " + " | ".join(lbl for _, lbl in KEYS) + " |")
    print("|---|" + "---:|" * len(KEYS))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        cells = []
        for k, _ in KEYS:
            cells.append(f"{r[idx[k]]} {units[idx[k]]}".strip() if k in idx else "n/a")
        print(f"| `{name}` grid {r[idx['Grid Size']]} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        fl = sys.argv[sys.argv.index("--from-last") + 1] if "--from-last" in sys.argv else None
        launches(sys.argv[2], fl)
    else:
        full(sys.argv[2])
