"""Extracts the metrics we track from an .ncu-rep into a small text summary (run on the CPU box)."""
import csv
import io
import re
import subprocess
import sys

KEYS = [
    r"^gpu__time_duration\.sum$", r"^sm__cycles_elapsed\.avg$", r"^sm__cycles_elapsed\.avg\.per_second$",
    r"^launch__registers_per_thread$", r"^launch__grid_size$", r"^launch__block_size$",
    r"^dram__bytes_read\.sum$", r"^dram__bytes_write\.sum$", r"^dram__bytes_read\.sum\.per_second$",
    r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"sm__pipe_tensor_cycles_active_realtime\.avg\.pct_of_peak_sustained_elapsed",
    r"sm__pipe_tensor_subpipe_imma_cycles_active_realtime\.avg$",
    r"^sm__inst_executed_pipe_tensor_subpipe_imma\.avg\.pct_of_peak_sustained_active$",
    r"^sm__mem_tensor_cycles_active\.avg\.pct_of_peak_sustained_elapsed$",
    r"^sm__inst_executed_pipe_alu\.avg\.pct_of_peak_sustained_active$",
    r"^sm__inst_executed_pipe_fma\.avg\.pct_of_peak_sustained_active$",
    r"^sm__inst_executed_pipe_tmem\.avg\.pct_of_peak_sustained_active$",
    r"^smsp__issue_active\.avg\.pct_of_peak_sustained_active$", r"^smsp__inst_executed\.sum$",
    r"^smsp__warps_active\.avg\.per_cycle_active$", r"^smsp__warps_eligible\.avg\.per_cycle_active$",
    r"^l1tex__m_xbar2l1tex_read_bytes\.sum$", r"^l1tex__m_xbar2l1tex_read_bytes\.sum\.per_second$",
    r"^lts__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^l1tex__data_pipe_tc_wavefronts_mem_shared\.sum\.pct_of_peak_sustained_elapsed$",
    r"^smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio$",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    pats = [re.compile(k) for k in KEYS]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"## {name[:110]}")
        for h, u, v in zip(hdr, units, r):
            if any(p.search(h) for p in pats) and v not in ("", "0"):
                print(f"{h} [{u}] = {v}")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
