#!/usr/bin/env python
"""Summarise an .ncu-rep (read here with `ncu -i`) into the handful of metrics DESIGN.md / profiles/ cite."""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("smsp__inst_executed.sum", "inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts.sum", "lsu_wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex_pct"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
    ("smsp__cycles_active.avg", "smsp_cycles"),
    ("sm__cycles_elapsed.max", "sm_cycles"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        print("==", r[ki][:90])
        parts = []
        for m, short in WANT:
            if m in hdr:
                i = hdr.index(m)
                parts.append("%s=%s%s" % (short, r[i], units[i] if units[i] not in ("", "%") else ("%" if units[i] else "")))
        print("   " + "  ".join(parts))


if __name__ == "__main__":
    main(sys.argv[1])
