"""Condenses `ncu --page raw --csv` exports (gpurun_out/<tag>_*_raw.csv, written by tools/run_profiles*.sh) into one CSV with the
columns the roofline discussion uses, normalised to microseconds / bytes, plus a markdown table.
Usage: python tools/ncu_summary.py <tag> <out_prefix>     e.g.  python tools/ncu_summary.py r2d profiles/r2d_ncu_full"""
import csv
import glob
import os
import sys

COLS = [("gpu__time_duration.sum", "time_us"), ("dram__bytes_read.sum", "dram_read_B"), ("dram__bytes_write.sum", "dram_write_B"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
        ("FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_peak"),
        ("lts__t_sectors_srcunit_tex_op_read.sum", "l2_read_sectors"), ("sm__cycles_active.avg", "sm_cycles_active"), ("sm__cycles_elapsed.avg", "sm_cycles_elapsed"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_scoreboard"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall_no_instruction"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_scoreboard"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait")]
SCALE = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3, "second": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def load(path, variant):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        rec = {"capture": variant, "kernel": r[idx["Kernel Name"]].split("(")[0]}
        for col, name in COLS:
            if col not in idx or r[idx[col]] == "":
                rec[name] = ""
                continue
            try:
                v = float(r[idx[col]].replace(",", ""))
            except ValueError:
                rec[name] = ""
                continue
            rec[name] = v * SCALE.get(units[idx[col]], 1.0)
        out.append(rec)
    return out


def main():
    tag, prefix = sys.argv[1], sys.argv[2]
    recs = []
    for p in sorted(glob.glob(os.path.join("gpurun_out", tag + "_*_raw.csv"))):
        recs += load(p, os.path.basename(p)[len(tag) + 1:-len("_raw.csv")])
    names = ["capture", "kernel"] + [n for _, n in COLS]
    with open(prefix + ".csv", "w", newline="") as f:
        w = csv.DictWriter(f, names)
        w.writeheader()
        for r in recs:
            w.writerow({k: (("%.6g" % v) if isinstance(v, float) else v) for k, v in r.items()})
    # one markdown row per (capture, kernel): mean over the captured launches
    agg = {}
    for r in recs:
        agg.setdefault((r["capture"], r["kernel"]), []).append(r)
    with open(prefix + ".md", "w") as f:
        f.write("| capture | kernel | launches | time us | DRAM read+write B | regs | grid x block | occupancy % | DRAM % of peak | L2 read sectors | long_scoreboard | barrier | no_instruction | SM cycles active / elapsed |\n")
        f.write("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for (cap, k), rs in sorted(agg.items()):
            def mean(n):
                v = [x[n] for x in rs if x[n] != ""]
                return sum(v) / len(v) if v else float("nan")
            f.write("| %s | %s | %d | %.1f | %.0f | %d | %d x %d | %.1f | %.2f | %.0f | %.1f | %.1f | %.1f | %.0f / %.0f |\n" % (
                cap, k, len(rs), mean("time_us"), mean("dram_read_B") + mean("dram_write_B"), mean("regs"), mean("grid"), mean("block"), mean("occupancy_pct"),
                mean("dram_pct_of_peak"), mean("l2_read_sectors"), mean("stall_long_scoreboard"), mean("stall_barrier"), mean("stall_no_instruction"),
                mean("sm_cycles_active"), mean("sm_cycles_elapsed")))
    print("wrote %s.csv / .md: %d launches, %d kernel rows" % (prefix, len(recs), len(agg)))


if __name__ == "__main__":
    main()
