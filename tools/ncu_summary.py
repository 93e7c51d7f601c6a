"""Print the handful of ncu metrics we steer by from a .ncu-rep (run where ncu is installed; no GPU needed).
usage: python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [more.ncu-rep ...]"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "launch__grid_size", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_membar_per_warp_active.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            rec = dict(zip(hdr, r))
            print(f"== {path}: {rec.get('Kernel Name', '?')[:90]}")
            for w in WANT:
                # (some metrics carry a section prefix in the raw page, e.g. "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime...")
                for col in hdr:
                    if (col == w or col.endswith("." + w)) and rec.get(col, "") != "":
                        print(f"  {w:<85}{rec[col]:>16} {units[hdr.index(col)]}")
                        break


if __name__ == "__main__":
    main()
