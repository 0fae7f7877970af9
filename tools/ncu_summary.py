#!/usr/bin/env python
"""Developer tool (runs here, no GPU): condense an .ncu-rep (ncu --set full) into the handful of
metrics DESIGN.md / bench.py quote, as a small CSV under profiles/.

    python tools/ncu_summary.py gpurun_out/r01a_render.ncu-rep profiles/r01a_render_summary.csv
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum",
    "sm__inst_executed_pipe_lsu.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__cycles_elapsed.max",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "metric", "unit", "value"])
        for r in rows[2:]:
            name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
            for m in METRICS:
                if m in col:
                    w.writerow([name, m, units[col[m]], r[col[m]]])
            tr = tw = None
            try:
                conv = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                tr = float(r[col["dram__bytes_read.sum"]]) * conv[units[col["dram__bytes_read.sum"]]]
                tw = float(r[col["dram__bytes_write.sum"]]) * conv[units[col["dram__bytes_write.sum"]]]
                w.writerow([name, "traffic_bytes(read+write)", "byte", int(tr + tw)])
            except Exception:
                pass
    print(open(out).read())


if __name__ == "__main__":
    main()
