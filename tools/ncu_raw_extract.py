"""stdin: `ncu -i X.ncu-rep --page raw --csv`; prints the handful of per-launch numbers that go
into profiles/ (duration, DRAM bytes, DRAM / L2 throughput, occupancy) for the LAST kernel row."""
import csv
import sys

WANT = ('Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size')


def main():
  rows = [r for r in csv.reader(sys.stdin) if r]
  # ncu prints "==PROF==" banner lines before the table when stderr is merged: keep from the header on
  start = next((i for i, r in enumerate(rows) if 'Kernel Name' in r), None)
  if start is None or len(rows) < start + 3:
    print('  no capture')
    return
  hdr, units, vals = rows[start], rows[start + 1], rows[-1]
  for h, u, v in zip(hdr, units, vals):
    if h in WANT:
      print('  %-66s %s %s' % (h, v, u))


if __name__ == '__main__':
  main()
