"""Summarise ncu output into markdown for profiles/:
   python tools/summarize_ncu.py full <report.ncu-rep>      -> per-launch table of a `--set full` capture
   python tools/summarize_ncu.py launches <launches.csv>    -> per-kernel-class share of a gpu__time_duration launch list
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

mode, path = sys.argv[1], sys.argv[2]

if mode == "full":
    # `path`: a .ncu-rep, or the `ncu -i <rep> --page raw --csv` export made on the GPU box (the reports exceed gpurun's 64 MiB)
    raw = open(path).read() if path.endswith(".csv") else \
        subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = OrderedDict([
        ("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
    ])
    idx = [(hdr.index(k), v) for k, v in cols.items() if k in hdr]
    print("| " + " | ".join(f"{v} [{units[i]}]" if units[i] else v for i, v in idx) + " |")
    print("|" + "---|" * len(idx))
    for r in data:
        cells = []
        for i, v in idx:
            x = r[i]
            if v == "kernel":
                x = x.replace("w2v2::", "").split("(")[0][:60]
            else:
                try:
                    x = f"{float(x):.3f}".rstrip("0").rstrip(".")
                except ValueError:
                    pass
            cells.append(x)
        print("| " + " | ".join(cells) + " |")
else:
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, cnt = {}, {}
    for r in rows[h + 1:]:
        name = r[ik].replace("w2v2::", "").split("(")[0]
        val = float(r[iv].replace(",", ""))
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)
        agg[name] = agg.get(name, 0.0) + val * scale
        cnt[name] = cnt.get(name, 0) + 1
    tot = sum(agg.values())
    print("| kernel | launches | total us | share |")
    print("|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
        print(f"| {k[:70]} | {cnt[k]} | {v:.1f} | {100 * v / tot:.1f} % |")
    print(f"| **all** | {sum(cnt.values())} | {tot:.1f} | 100 % |")
