"""Turns ncu reports (gpurun_out/*.ncu-rep) and launch lists into the small text summaries kept under profiles/.
Usage: python scripts/ncu_summary.py report gpurun_out/x.ncu-rep profiles/x.txt
       python scripts/ncu_summary.py launches gpurun_out/launches.csv profiles/launches_summary.txt"""
import csv, io, re, subprocess, sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.per_cycle_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum"]


def ncu(args):
    return subprocess.run(["ncu"] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def report(rep, out):
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    head, units, rows = raw[0], raw[1], raw[2:]
    lines = ["ncu --set full summary of %s" % rep, ""]
    for r in rows:
        d = dict(zip(head, r))
        u = dict(zip(head, units))
        lines.append("kernel: %s" % d.get("Kernel Name", "?"))
        for k in KEYS:
            for h in head:
                if h == k or h.endswith("." + k):
                    if d.get(h, "") != "":
                        lines.append("   %-82s %s %s" % (k, d[h], u.get(h, "")))
                    break
        lines.append("")
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv"]))))
    # stall reasons aggregated over the kernel(s) in the report (warp sampling)
    tot, n = defaultdict(int), 0
    hdr = None
    for r in src:
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr) or not r[0].startswith("0x"):
            continue
        d = dict(zip(hdr, r))
        n += int(d.get("# Samples") or 0)
        for k, v in d.items():
            if k.startswith("stall_") and "Not Issued" not in k and v:
                tot[k] += int(v)
    if n:
        lines.append("warp-state samples (all kernels in the report): %d" % n)
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]:
            lines.append("   %-24s %5.1f %%" % (k, 100.0 * v / n))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


def launches(path, out):
    rows = list(csv.reader(open(path, errors="replace")))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    head = rows[start]
    ci = {c: i for i, c in enumerate(head)}
    agg = defaultdict(lambda: [0, 0.0])
    total = 0.0
    for r in rows[start + 1:]:
        if len(r) != len(head) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[ci["Kernel Name"]]).replace("void ", "").replace("slmm::", "")
        v = float(r[ci["Metric Value"]].replace(",", ""))
        unit = r[ci["Metric Unit"]]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        agg[name][0] += 1
        agg[name][1] += us
        total += us
    lines = ["launch list summary of %s (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised:"
             " compare shares)" % path, "%-70s %8s %12s %7s" % ("kernel", "launches", "total us", "share")]
    for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("%-70s %8d %12.1f %6.1f%%" % (name[:70], cnt, us, 100.0 * us / total))
    lines.append("%-70s %8d %12.1f" % ("TOTAL", sum(v[0] for v in agg.values()), total))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    {"report": report, "launches": launches}[sys.argv[1]](sys.argv[2], sys.argv[3])
