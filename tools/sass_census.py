"""Census of the Blackwell-specific SASS mnemonics per kernel of the built objects (what proves the
tcgen05 / TMA path, /opt/skills/guides/B200_PROFILING.md): UTCHMMA (tcgen05.mma, by kind), UTMALDG
(TMA tensor loads), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit), SYNCS (mbarrier), F2FP (the bf16
operand split), plus anything that would mean a non-Blackwell path (HMMA, wgmma).
  python tools/sass_census.py sg2im_b200/csrc/conv_tc.o sg2im_b200/csrc/conv_wgrad_tc.o > profiles/r02_sass_census.txt
"""
import collections
import re
import subprocess
import sys

WANT = ('UTCHMMA', 'UTCQMMA', 'UTCIMMA', 'UTMALDG', 'UTMASTG', 'UTMAPF', 'LDTM', 'STTM', 'UTCBAR', 'UTCATOMSWS',
        'SYNCS', 'F2FP', 'FENCE', 'HMMA', 'WGMMA', 'REDG', 'ELECT')


def main():
  for path in sys.argv[1:]:
    out = subprocess.run(['cuobjdump', '-sass', path], capture_output=True, text=True, check=True).stdout
    name, per = None, collections.OrderedDict()
    for line in out.splitlines():
      m = re.search(r'Function : (\S+)', line)
      if m:
        dem = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r'_GLOBAL__N__[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]+', 'anon', dem)
        name = re.sub(r'\(anonymous namespace\)::|<unnamed>::', '', dem).split('(')[0]
        per[name] = collections.Counter()
        continue
      m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
      if m and name is not None:
        op = m.group(1)
        for w in WANT:
          if op.startswith(w):
            per[name]['.'.join(op.split('.')[:3])] += 1
    print('== %s' % path)
    for k, c in per.items():
      if not c:
        continue
      print('%s' % k[:120])
      print('    ' + '  '.join('%s x%d' % kv for kv in sorted(c.items())))
  return 0


if __name__ == '__main__':
  sys.exit(main())
