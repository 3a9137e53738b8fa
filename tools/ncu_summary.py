"""Key counters of an `ncu --set full` report (run where ncu is installed; no GPU needed):
python tools/ncu_summary.py gpurun_out/r01_<kernel>.ncu-rep"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'lts__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio']


def main(path):
  out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
  rows = list(csv.reader(out.splitlines()))
  hdr, units = rows[0], rows[1]
  idx = {h: i for i, h in enumerate(hdr)}
  print('# %s' % path)
  for r in rows[2:]:
    print('launch %s  %s' % (r[idx['ID']], r[idx['Kernel Name']][:100]))
    for w in WANT:
      if w in idx:
        print('   %-82s %s %s' % (w, r[idx[w]], units[idx[w]]))


if __name__ == '__main__':
  for p in sys.argv[1:]:
    main(p)
