#!/usr/bin/env python3
"""Summarise an `ncu --set full --import-source on` report into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/X_prof.ncu-rep [images_per_launch] > profiles/X_ncu_summary.txt

Prints the headline raw metrics (duration, DRAM bytes, tensor/IMMA pipe, issue utilisation, registers, smem
conflicts, stall breakdown) and the SASS opcode mix with warp-stall sample shares from the source page.
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "lts__t_bytes.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__warps_active.avg.per_cycle_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    n_img = int(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = ncu_csv(rep, "raw")
    hdr, units = raw[0], raw[1]
    print(f"# ncu summary of {rep.split('/')[-1]} ({len(raw) - 2} profiled launch(es)); values of launch 0")
    vals = raw[2]
    d = dict(zip(hdr, zip(units, vals)))
    print("kernel:", d.get("Kernel Name", ("", "?"))[1], " grid", d.get("Grid Size", ("", "?"))[1], " block", d.get("Block Size", ("", "?"))[1])
    for k in KEYS:
        if k in d:
            print(f"{k:90s} {d[k][1]:>16s} {d[k][0]}")
    stalls = [(h, float(v)) for h, v in zip(hdr, vals) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    print("\n# warp stall reasons (warps stalled per issue-active cycle)")
    for h, v in sorted(stalls, key=lambda x: -x[1])[:8]:
        print(f"{h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {v:8.3f}")
    if n_img:
        t = float(d["gpu__time_duration.sum"][1]) * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}[d["gpu__time_duration.sum"][0]]
        rd = float(d["dram__bytes_read.sum"][1]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[d["dram__bytes_read.sum"][0]]
        wr = float(d["dram__bytes_write.sum"][1]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[d["dram__bytes_write.sum"][0]]
        cyc = float(d["sm__cycles_elapsed.avg"][1])
        sms = 148
        print(f"\n# derived for {n_img} images per launch")
        print(f"images/s under ncu (cold, serialised): {n_img / t:,.0f}")
        print(f"DRAM traffic per launch: {rd + wr:,.0f} B = {(rd + wr) / n_img:,.0f} B/image (algorithmic 32768 B/image for the conv stack "
              f"with features out, 16428 B/image predictions only; read {rd / n_img:,.0f} + written {wr / n_img:,.0f})")
        print(f"SM cycles per image per SM: {cyc / (n_img / sms):,.0f}")
    src = ncu_csv(rep, "source")
    if len(src) > 3:
        h = src[1]
        ia, ie, isamp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
        ops, samp = collections.Counter(), collections.Counter()
        for r in src[2:]:
            if len(r) <= max(ia, ie, isamp):
                continue
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ia])
            op = m.group(2).split(".")[0] if m else "?"
            ops[op] += int(r[ie] or 0)
            samp[op] += int(r[isamp] or 0)
        tot, ts = sum(ops.values()), max(1, sum(samp.values()))
        print(f"\n# SASS opcode mix (warp-level instructions executed: {tot:,}; stall samples: {ts:,})")
        for op, c in ops.most_common():              # the whole table: UTCIMMA / LDTM / UTMALDG are rare but are the point
            per = f"{c / n_img:10.1f}/image" if n_img else ""
            print(f"{op:12s} {c:14,d} {100 * c / tot:5.1f}% {per}   samples {100 * samp[op] / ts:5.1f}%")


if __name__ == "__main__":
    main()
