"""Text summary of ncu --set full captures (.ncu-rep) for profiles/: duration, DRAM bytes, occupancy
limits, issue utilisation, top stall reasons, tensor-pipe activity (must be zero on this path).
Usage: python tools/summarize_ncu.py gpurun_out/r01i_full_*.ncu-rep > profiles/r01i_ncu_full_summary.txt"""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum"]


def main():
    print("# ncu --set full --clock-control none --import-source on, one launch each; values from 'ncu -i <rep> --page raw --csv'")
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        if len(rows) < 3:
            print("== %s: no kernel captured" % rep)
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            print("== %s" % rep.split("/")[-1])
            for w in WANT:
                if w in hdr:
                    print("  %-72s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
            stalls = [(float(r[i].replace(",", "")), h) for i, h in enumerate(hdr)
                      if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and r[i]]
            for x, h in sorted(stalls, reverse=True)[:6]:
                name = h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")
                print("  stall %-30s %6.2f warps per issue-active cycle" % (name, x))


if __name__ == "__main__":
    main()
