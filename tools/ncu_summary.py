"""Summarise an .ncu-rep (read here, no GPU needed): one line per captured launch with the metrics DESIGN.md quotes.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [more.ncu-rep ...] > profiles/<name>_summary.txt
"""
import csv
import io
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "us",
    "dram__bytes_read.sum": "dram_rd",
    "dram__bytes_write.sum": "dram_wr",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "launch__registers_per_thread": "regs",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_pct",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__cluster_dim_x": "cluster",
    "sm__inst_issued.avg.pct_of_peak_sustained_active": "issue_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pct",
    "sm__cycles_elapsed.avg.per_second": "sm_hz",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_sb",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_sb",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio": "stall_no_inst",
    "launch__shared_mem_per_block_dynamic": "smem_dyn",
    "launch__occupancy_limit_registers": "occ_lim_regs",
}


def main():
    for path in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        ki = hdr.index("Kernel Name")
        print(f"# {path}")
        for r in rows[2:]:
            d = {"kernel": r[ki][:70]}
            for h, u, v in zip(hdr, units, r):
                if h in KEYS:
                    d[KEYS[h]] = f"{v} {u}".strip()
            print(d)


if __name__ == "__main__":
    main()
