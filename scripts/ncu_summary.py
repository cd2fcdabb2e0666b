"""Summarise an .ncu-rep (read here, no GPU needed) into the small CSVs kept under profiles/:
   python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/name.csv
writes `metric,unit,value` for the headline counters plus, from the source page, the executed
thread-instruction mix per unit and the stall-sample shares per opcode class."""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
    "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__icc_request_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct",
]


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    units = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    raw = ncu(rep, "raw")
    hdr, unit, val = raw[0], raw[1], raw[2]
    d = {h: (u, v) for h, u, v in zip(hdr, unit, val)}
    lines = [("metric", "unit", "value"), ("Kernel Name", "", d.get("Kernel Name", ("", ""))[1])]
    for k in KEYS:
        if k in d:
            lines.append((k, d[k][0], d[k][1]))
    for h in hdr:
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(d[h][1] or 0) >= 0.01:
            lines.append((h, d[h][0], d[h][1]))
    src = ncu(rep, "source")
    if len(src) > 2:
        sh = src[1]
        ix = {h: i for i, h in enumerate(sh)}
        cnt, stall = collections.Counter(), collections.defaultdict(collections.Counter)
        names = [h for h in sh if h.startswith("stall_") and "Not Issued" not in h]
        for r in src[2:]:
            op = re.sub(r"^@!?U?P\d+\s+", "", r[ix["Source"]].strip()).split()[0]
            cls = "IMAD.WIDE" if op.startswith("IMAD.WIDE") else ("IMAD.other" if op.startswith("IMAD") else op.split(".")[0])
            cnt[cls] += int(r[ix["Thread Instructions Executed"]])
            for s in names:
                stall[cls][s] += int(r[ix[s]])
        tot = sum(cnt.values())
        allst = sum(sum(c.values()) for c in stall.values())
        for cls, c in cnt.most_common(10):
            per = f"{c / units:.0f}" if units else ""
            lines.append((f"thread_inst_executed[{cls}]", "share / per unit", f"{c / tot:.4f} / {per}"))
        if units:
            lines.append(("thread_inst_executed[total]", "per unit", f"{tot / units:.0f}"))
        agg = collections.Counter()
        for c in stall.values():
            agg.update(c)
        for s, v in agg.most_common(7):
            lines.append((f"stall_samples[{s}]", "share", f"{v / allst:.4f}"))
        for cls in ("IMAD.WIDE", "IADD3"):
            t = sum(stall[cls].values()) or 1
            top = ", ".join(f"{s}={v / t:.2f}" for s, v in stall[cls].most_common(4))
            lines.append((f"stall_samples_at[{cls}]", "share of its samples", top))
    with open(dst, "w", newline="") as f:
        csv.writer(f).writerows(lines)
    print(open(dst).read())


if __name__ == "__main__":
    main()
