"""Pick the metrics that matter for these kernels out of `ncu -i X.ncu-rep --page raw --csv` output."""
import csv
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__cycles_elapsed.max',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg',
        'SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg',
        'SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_shared.avg',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_red.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lts__t_sectors.sum', 'lts__t_sectors_op_red.sum',
        'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'lts__t_bytes.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_drain_per_issue_active.ratio']


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== kernel:", r[hdr.index("Kernel Name")][:60] if "Kernel Name" in hdr else "")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k} [{units[i]}] {r[i]}")


if __name__ == "__main__":
    main()
