"""Condense an `ncu --set full` report into the handful of numbers DESIGN.md and profiles/ncu_traffic.json quote.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [more.ncu-rep ...] > profiles/r2x_ncu_summary.csv
One row per profiled launch: kernel, grid, block, duration, dram bytes read / written, L2 bytes, hit rate, registers,
achieved occupancy, issue-slot utilisation, NVLink bytes where the counters exist."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "nvltx__bytes.sum", "nvlrx__bytes.sum", "lts__t_sectors_srcunit_ltcfabric.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]


def main():
    w = csv.writer(sys.stdout)
    w.writerow(["report", "kernel"] + WANT)
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        if len(rows) < 3:
            continue
        head, units = rows[0], rows[1]
        col = {c: i for i, c in enumerate(head)}
        for r in rows[2:]:
            vals = []
            for m in WANT:
                vals.append(f"{r[col[m]]} {units[col[m]]}".strip() if m in col else "")
            w.writerow([rep.rsplit("/", 1)[-1], r[col["Kernel Name"]][:60]] + vals)


if __name__ == "__main__":
    main()
