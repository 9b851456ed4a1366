"""Summarise an `ncu --set full` report (one row per captured kernel) into JSON: the metrics DESIGN.md quotes.
usage: ncu_summary.py <report.ncu-rep> <out.json> [units_per_launch] [note]
Runs `ncu -i <report> --page raw --csv` (ncu only reads the file here; nothing is profiled)."""
import csv, io, json, subprocess, sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block", "launch__grid_size", "launch__block_size",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "sm__cycles_active.min", "sm__cycles_active.max", "sm__cycles_active.avg"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    units = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    note = sys.argv[4] if len(sys.argv) > 4 else ""
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    head, unit_row, data = rows[0], rows[1], rows[2:]
    kernels = []
    for r in data:
        k = {}
        for key in KEYS:
            if key in head:
                i = head.index(key)
                k[key] = r[i] if key == "Kernel Name" else f"{r[i]} {unit_row[i]}".strip()
        if units and "smsp__inst_executed.sum" in head:      # warp instructions x 32 lanes (an upper bound of the thread count)
            k["thread_instructions_per_unit"] = round(32.0 * float(r[head.index("smsp__inst_executed.sum")].replace(",", "")) / units, 1)
        kernels.append(k)
    json.dump({"source": note, "units_per_launch": units, "kernels": kernels}, open(out, "w"), indent=1)
    for k in kernels:
        print(k["Kernel Name"], k.get("gpu__time_duration.sum"), k.get("thread_instructions_per_unit"))


if __name__ == "__main__":
    main()
