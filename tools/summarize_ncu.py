"""Summarise ncu outputs brought back from gpurun into small committed text files under profiles/.
  python tools/summarize_ncu.py <tag>     (expects gpurun_out/launches_<tag>.csv and gpurun_out/prof_<tag>.ncu-rep)"""
import csv
import io
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)


def short(name):
    name = re.sub(r"sb200::\(anonymous namespace\)::", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name[:90]


lp = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(lp):
    rows = []
    with open(lp) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    for r in rd:
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            if unit in ("usecond", "us"):
                v *= 1e3
            elif unit in ("msecond", "ms"):
                v *= 1e6
            rows.append((short(r["Kernel Name"]), v))
    agg = OrderedDict()
    for k, v in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v for _, v in rows)
    with open(os.path.join(out_dir, f"{tag}_launches_summary.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, first {len(rows)} launches of "
                f"`python bench.py --steps 1 --warmup 3` (cold-cache, serialised: compare SHARES)\n")
        f.write(f"# total {tot / 1e6:.3f} ms over {len(rows)} launches\n")
        f.write(f"{'kernel':92s} {'launches':>8s} {'total_ms':>10s} {'avg_us':>10s} {'share':>7s}\n")
        for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:92s} {n:8d} {v / 1e6:10.3f} {v / n / 1e3:10.2f} {100 * v / tot:6.2f}%\n")
    print(open(os.path.join(out_dir, f"{tag}_launches_summary.txt")).read())

rp = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
if os.path.exists(rp):
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rd[0], rd[1], rd[2:]
    want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tensor.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
            "smsp__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum", "lts__t_bytes.sum"]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(os.path.join(out_dir, f"{tag}_ncu_full_summary.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on; {len(data)} captured launches\n")
        for row in data:
            f.write("-" * 100 + "\n")
            for w in want:
                if w in idx:
                    f.write(f"{w:75s} {row[idx[w]]:>20s} {units[idx[w]]}\n")
    print(open(os.path.join(out_dir, f"{tag}_ncu_full_summary.txt")).read())
