"""Summary of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv
--log-file X.csv ...`): launches and total time per kernel (template arguments kept, argument lists
dropped), share of the kernel time, ATen share.  Per-launch times under ncu are cold-cache and
serialised: the shares are what carries over to the un-profiled step.
Usage: python tools/ncu_launch_summary.py X.csv "header line" > profiles/X_summary.txt"""
import collections
import csv
import re
import sys


def short(name):
  name = re.sub(r'\(anonymous namespace\)::|<unnamed>::', '', name)
  name = re.sub(r'^void ', '', name)
  depth, out = 0, []
  for ch in name:                       # cut at the first '(' outside template brackets
    if ch == '<':
      depth += 1
    elif ch == '>':
      depth -= 1
    elif ch == '(' and depth == 0:
      break
    out.append(ch)
  return ''.join(out)[:94]


def main():
  rows = [r for r in csv.reader(open(sys.argv[1], errors='replace')) if len(r) > 14 and r[0].isdigit()]
  tot, cnt = collections.Counter(), collections.Counter()
  for r in rows:
    unit, val = r[13], float(r[14].replace(',', ''))
    us = val * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}.get(unit, 1e-3)
    k = short(r[4])
    tot[k] += us
    cnt[k] += 1
  total = sum(tot.values())
  aten = sum(v for k, v in tot.items() if k.startswith('at::'))
  if len(sys.argv) > 2:
    print(sys.argv[2])
  print('%d launches, total kernel time %.1f us; ATen (at::) kernels: %d launches, %.1f us = %.1f %% of kernel time' % (
      len(rows), total, sum(c for k, c in cnt.items() if k.startswith('at::')), aten, 100 * aten / max(total, 1e-9)))
  for k, v in tot.most_common(60):
    print('%-96s %6d %12.1f us %6.1f%%' % (k, cnt[k], v, 100 * v / total))


if __name__ == '__main__':
  main()
