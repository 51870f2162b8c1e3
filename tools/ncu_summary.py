#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): one line per kernel launch with the metrics the
roofline discussion needs.   python tools/ncu_summary.py gpurun_out/x.ncu-rep [--stalls]"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "us"),
    ("dram__bytes_read.sum", "rdMB"),
    ("dram__bytes_write.sum", "wrMB"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__t_sectors_op_read.sum", "l2rdsec"),
    ("smsp__inst_executed.sum", "winst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smemwf"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    name_i = idx.get("Kernel Name")
    for r in rows[2:]:
        parts = [r[name_i][:46]]
        for k, short in KEYS:
            if k in idx:
                v = r[idx[k]]
                u = units[idx[k]]
                try:
                    f = float(v.replace(",", ""))
                    if short in ("rdMB", "wrMB"):
                        f = f * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)
                    if short == "us":
                        f = f * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
                    v = f"{f:.4g}"
                except ValueError:
                    pass
                parts.append(f"{short}={v}")
        print("  ".join(parts))
        if "--stalls" in sys.argv:
            st = []
            for h, i in idx.items():
                if h.startswith(STALL) and h.endswith("_per_issue_active.ratio"):
                    try:
                        st.append((float(r[i]), h[len(STALL):-len("_per_issue_active.ratio")]))
                    except ValueError:
                        pass
            st.sort(reverse=True)
            print("     stalls/issue: " + "  ".join(f"{n}={v:.2f}" for v, n in st[:7]))


if __name__ == "__main__":
    main()
