"""Per-kernel SASS comparison of two object files / libraries (sm_100a): prints, for every kernel
present in both, whether the instruction streams are identical (addresses stripped; the kernels are
matched by their demangled names without the per-translation-unit hash of the anonymous
namespace).  Used to show that adding an opt-in kernel variant leaves the validated default
kernels byte-identical:
    python tools/sass_compare.py <old.o|.so> <new.o|.so> [name-substring ...]
"""
import hashlib
import re
import subprocess
import sys


def kernels(path):
  out = subprocess.run(['cuobjdump', '-sass', path], capture_output=True, text=True, check=True).stdout
  res, name = {}, None
  for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
      dem = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
      name = re.sub(r'_GLOBAL__N__[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]+', 'anon', dem)
      name = re.sub(r'\(anonymous namespace\)|<unnamed>', 'anon', name)
      res[name] = []
      continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(.*?)\s*/\*', line)
    if m and name is not None:
      res[name].append(re.sub(r'\s+', ' ', m.group(1)))
  return {k: (len(v), hashlib.md5('\n'.join(v).encode()).hexdigest()[:10]) for k, v in res.items()}


def main():
  a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
  want = sys.argv[3:]
  same = diff = 0
  for k in sorted(set(a) & set(b)):
    if want and not any(w in k for w in want):
      continue
    ok = a[k] == b[k]
    same += ok
    diff += not ok
    print('%-9s %6d instr %s  %s' % ('identical' if ok else 'DIFFERENT', b[k][0], b[k][1], k[:110]))
  print('only in new: %d kernels' % len(set(b) - set(a)))
  for k in sorted(set(b) - set(a)):
    print('   new       %6d instr %s  %s' % (b[k][0], b[k][1], k[:110]))
  print('%d identical, %d different' % (same, diff))
  return 1 if diff else 0


if __name__ == '__main__':
  sys.exit(main())
