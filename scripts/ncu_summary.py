#!/usr/bin/env python
"""Turns an .ncu-rep (ncu --set full --import-source on) into the text summary kept under profiles/.

    python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_summary.txt

Per kernel: duration, launch shape, occupancy limits, issue rate, pipe utilisation, DRAM bytes, L1/L2 hit
rates, the warp-stall breakdown (sampling), the executed-instruction mix and the hottest stall sites."""
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.per_cycle_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    rows = ncu(rep, "raw")
    hdr, units = rows[0], rows[1]
    print("# %s" % rep)
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("\n== %s" % name)
        for h, u, v in zip(hdr, units, r):
            if h in KEYS:
                print("   %-62s %-14s %s" % (h, u, v))
        print("   warp stalls per issued instruction (cycles):")
        st = [(h, float(v)) for h, v in zip(hdr, r) if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and v]
        for h, v in sorted(st, key=lambda kv: -kv[1]):
            if v >= 0.05:
                print("      %-28s %.2f" % (h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), v))
    src = ncu(rep, "source")
    blocks, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None and r:
            cur["rows"].append(r)
    seen = set()
    for bl in blocks:
        if bl["name"] in seen or len(bl["rows"]) < 2:
            continue
        seen.add(bl["name"])
        hdr, data = bl["rows"][0], bl["rows"][1:]
        ix = {h: i for i, h in enumerate(hdr)}
        if "# Samples" not in ix:
            continue
        tot = sum(int(r[ix["# Samples"]] or 0) for r in data) or 1
        exe = {}
        for r in data:
            op = re.sub(r"^@!?U?P\d+\s+", "", r[ix["Source"]].strip()).split()[0].split(".")[0] if r[ix["Source"]].strip() else "?"
            exe[op] = exe.get(op, 0) + int(r[ix["Instructions Executed"]] or 0)
        te = sum(exe.values()) or 1
        print("\n== %s : source view (%d SASS lines, %d stall samples)" % (bl["name"], len(data), tot))
        print("   executed instruction mix: " + ", ".join("%s %.1f%%" % (k, 100.0 * v / te)
                                                           for k, v in sorted(exe.items(), key=lambda kv: -kv[1])[:16]))
        reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        print("   hottest SASS lines (share of stall samples, top reasons):")
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:12]:
            s = int(r[ix["# Samples"]] or 0)
            rs = sorted(((h, int(r[ix[h]] or 0)) for h in reasons), key=lambda kv: -kv[1])[:2]
            print("      %5.2f%%  %-58s %s" % (100.0 * s / tot, r[ix["Source"]].strip()[:58], ", ".join("%s=%d" % kv for kv in rs)))


if __name__ == "__main__":
    main()
