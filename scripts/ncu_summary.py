"""Summarise an .ncu-rep (read offline, no GPU needed) into profiles/<name>.md: per-launch key metrics the roofline uses."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__waves_per_multiprocessor', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'SM_A.TriageCompute.sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed', 'smsp__pcsamp_warps_issue_stalled_long_scoreboard', 'smsp__pcsamp_warps_issue_stalled_selected',
        'smsp__pcsamp_warps_issue_stalled_wait', 'smsp__pcsamp_warps_issue_stalled_barrier', 'smsp__pcsamp_warps_issue_stalled_math_pipe_throttle', 'smsp__pcsamp_warps_issue_stalled_short_scoreboard',
        'smsp__pcsamp_warps_issue_stalled_no_instructions', 'smsp__pcsamp_warps_issue_stalled_lg_throttle', 'smsp__pcsamp_sample_count',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_red.sum', 'l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum']


def main(rep, out, title):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, 'w') as fh:
        fh.write(f'# {title}\n\nsource: `{rep}` (ncu --set full --clock-control none), read offline with `ncu -i ... --page raw --csv`\n\n')
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            fh.write(f"## launch id {d.get('ID')}: `{d.get('Kernel Name', '')[:90]}`  grid {d.get('Grid Size')} block {d.get('Block Size')}\n\n| metric | unit | value |\n|---|---|---|\n")
            for k in KEYS:
                if k in d and d[k] != '':
                    fh.write(f'| {k} | {units[hdr.index(k)]} | {d[k]} |\n')
            fh.write('\n')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else 'ncu summary')
