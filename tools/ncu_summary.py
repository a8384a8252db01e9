"""Summarises ncu captures brought back in gpurun_out/ into small text files under profiles/ (tracked)."""
import csv
import subprocess
import sys
from collections import OrderedDict

METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def full(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none, source: {rep}\n")
        for r in rows[2:]:
            f.write(f"kernel: {r[hdr.index('Kernel Name')][:100]}\n")
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write(f"  {m} = {r[i]} {units[i]}\n")


def launches(path, out):
    agg = OrderedDict()
    with open(path) as f:
        rows = [r for r in csv.reader(f) if len(r) > 5]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    iu = hdr.index("Metric Unit")
    for r in rows[1:]:
        name = r[ik].split("(")[0]
        t = float(r[iv].replace(",", ""))
        t *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, source: {path}\n")
        f.write("# cold-cache, serialised launches: compare SHARES, not absolutes\n")
        f.write(f"{'kernel':60s} {'launches':>8s} {'total_ms':>12s} {'share':>8s}\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:60]:60s} {n:8d} {t:12.3f} {t / tot:8.3f}\n")


if __name__ == "__main__":
    kind, src, dst = sys.argv[1:4]
    {"full": full, "launches": launches}[kind](src, dst)
