#!/usr/bin/env python
"""Summarise a .ncu-rep (ncu --set full) into a small text file for profiles/: per captured kernel the duration,
DRAM bytes/throughput, L2 hit rate, occupancy, registers, top warp-stall reasons, and the hottest source lines.

    python tools/summarize_ncu.py gpurun_out/x.ncu-rep > profiles/x_summary.txt
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "smsp__cycles_active.avg",
        "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    rep = sys.argv[1]
    hdr, units, rows = raw(rep)
    name_i = hdr.index("Kernel Name")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    if not stall_cols:
        stall_cols = [i for i, h in enumerate(hdr) if "warp_issue_stalled" in h and h.endswith(".pct")]
    for r in rows:
        print("=" * 100)
        print(r[name_i][:160])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:70s} {r[i]:>16s} {units[i]}")
        st = []
        for i in stall_cols:
            try:
                st.append((float(r[i]), hdr[i]))
            except ValueError:
                pass
        st.sort(reverse=True)
        print("  top stall reasons:")
        for v, h in st[:7]:
            print(f"    {v:10.3f}  {h}")
    # hottest source lines (CUDA-C view): aggregated over the captured launches of each function
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    agg = {}
    fn = fpath = None
    cols = None
    for row in csv.reader(io.StringIO(src)):
        if not row:
            continue
        if row[0] == "File Path":
            fpath = row[1]
            continue
        if row[0] == "Function Name":
            fn = row[1][:90]
            continue
        if row[0] == "Line No":
            cols = row
            continue
        if cols is None or fpath is None or "/root/repo/" not in fpath:
            continue
        try:
            si = cols.index("Warp Stall Sampling (All Samples)")
            ii = cols.index("Instructions Executed")
            samples, insts = float(row[si]), float(row[ii])
        except (ValueError, IndexError):
            continue
        key = (fn, fpath.split("/")[-1], row[0])
        a = agg.setdefault(key, [0.0, 0.0, row[1].strip()[:100]])
        a[0] += samples
        a[1] += insts
    by_fn = {}
    for (f, file, line), (sm, ins, text) in agg.items():
        by_fn.setdefault(f, []).append((sm, ins, file, line, text))
    for f, lst in by_fn.items():
        tot_s = sum(x[0] for x in lst) or 1.0
        tot_i = sum(x[1] for x in lst) or 1.0
        print("=" * 100)
        print("hot source lines (share of warp-stall samples | share of executed instructions):", f)
        for sm, ins, file, line, text in sorted(lst, reverse=True)[:16]:
            print(f"  {100 * sm / tot_s:5.1f}% | {100 * ins / tot_i:5.1f}%  {file}:{line:>4s}  {text}")


if __name__ == "__main__":
    main()
