"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) and a --set full report.
usage: python profiles/summarize.py launches.csv [report.ncu-rep]"""
import collections
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) > vi:
            agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
    mine = {k: v for k, v in agg.items() if "yc::" in k and "pack" not in k}
    tot = sum(sum(v) / len(v) for v in mine.values())
    print(f"{'kernel':72s} {'n':>4s} {'avg us':>9s} {'share':>7s}")
    for k, v in agg.items():
        avg = sum(v) / len(v) / 1000
        share = f"{100 * avg * 1000 / tot:6.1f}%" if k in mine else "      -"
        print(f"{k[:72]:72s} {len(v):4d} {avg:9.2f} {share}")
    print(f"step total (yc:: kernels, per step): {tot / 1000:.2f} us")


def report(path):
    out = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(out.splitlines()))
    h = rows[0]
    kn = h.index("Kernel Name")
    for r in rows[2:]:
        print("kernel:", r[kn][:80])
        for i, name in enumerate(h):
            if name in WANT:
                print(f"  {name:70s} {r[i]:>18s} {rows[1][i]}")


if __name__ == "__main__":
    launches(sys.argv[1])
    if len(sys.argv) > 2:
        report(sys.argv[2])
