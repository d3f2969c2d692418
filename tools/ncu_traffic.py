"""profiles/traffic.json from `ncu --set full` captures: for each arithmetic, the DRAM bytes
(dram__bytes_read.sum + dram__bytes_write.sum) of ONE launch of the step's top kernel, next to the
algorithmic bytes of that launch.  bench.py quotes it as `roofline.traffic`.

  python tools/ncu_traffic.py bf16x3=gpurun_out/r02_prof_halo_bf16x3.ncu-rep tf32=... \
      --kernel conv_tc_halo_kernel --algorithmic 738197504 --out profiles/traffic.json
"""
import argparse
import csv
import io
import json
import subprocess


def raw_rows(rep):
  out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True,
                       check=True).stdout
  rows = [r for r in csv.reader(io.StringIO(out)) if r]
  start = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
  hdr, units = rows[start], rows[start + 1]
  return [dict(zip(hdr, r)) for r in rows[start + 2:]], dict(zip(hdr, units))


def to_bytes(v, unit):
  v = float(v.replace(',', ''))
  return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('reps', nargs='+', help='math=path.ncu-rep')
  ap.add_argument('--kernel', required=True)
  ap.add_argument('--algorithmic', type=float, required=True)
  ap.add_argument('--out', default='profiles/traffic.json')
  a = ap.parse_args()
  res = {}
  for item in a.reps:
    math, rep = item.split('=', 1)
    rows, units = raw_rows(rep)
    rows = [r for r in rows if a.kernel in r['Kernel Name']]
    r = rows[-1]
    rd = to_bytes(r['dram__bytes_read.sum'], units['dram__bytes_read.sum'])
    wr = to_bytes(r['dram__bytes_write.sum'], units['dram__bytes_write.sum'])
    res[math] = {'kernel': r['Kernel Name'][:80], 'bytes_per_launch': rd + wr, 'dram_read': rd,
                 'dram_write': wr, 'algorithmic_bytes': a.algorithmic, 'source': rep.split('/')[-1]}
  json.dump(res, open(a.out, 'w'), indent=1)
  print(json.dumps(res, indent=1))


if __name__ == '__main__':
  main()
