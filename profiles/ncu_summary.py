"""Short text summary of the first kernel in an .ncu-rep (raw page), for profiles/*.txt:

    python profiles/ncu_summary.py gpurun_out/NAME.ncu-rep > profiles/NAME.txt
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "launch__grid_size",
        "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:3]:
        d = dict(zip(hdr, zip(units, vals)))
        print(f"{rep}\nkernel: {d.get('Kernel Name', ('', '?'))[1]}")
        for k in KEYS:
            if k in d:
                print(f"  {k:80s} {d[k][1]} {d[k][0]}")
        stalls = sorted(((float(v[1]), h) for h, v in d.items()
                         if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")),
                        reverse=True)[:6]
        print("  stalls per issue: " + ", ".join(f"{h.split('stalled_')[1].split('_per_')[0]} {x:.2f}" for x, h in stalls))


if __name__ == "__main__":
    main()
