"""Regenerate profiles/ptxas_summary.txt (registers / barriers / smem / spills per kernel, from the -Xptxas -v logs the
Makefile keeps under csrc/_build) and profiles/sass_evidence.txt (counts of the SASS mnemonics that prove the tcgen05 /
TMEM / TMA / mbarrier / cluster paths are what was compiled; /opt/skills/guides/B200_PROFILING.md lists them).
Run after `make` in rec-attend-public_b200/csrc; needs no GPU."""
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, 'rec-attend-public_b200', 'csrc', '_build')
SO = os.path.join(ROOT, 'rec-attend-public_b200', 'librecattend_b200.so')


def demangle(names):
  out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout.split('\n')
  res = []
  for n in out[:len(names)]:
    n = re.sub(r'\(anonymous namespace\)::', '', n)
    n = re.sub(r'^void ', '', n)
    n = re.sub(r'\(.*$', '', n)
    res.append(n)
  return res


def ptxas():
  rows = []
  for log in sorted(glob.glob(os.path.join(BUILD, '*.ptxas.log'))):
    unit = os.path.basename(log)[:-len('.ptxas.log')]
    lines = open(log).read().split('\n')
    i = 0
    while i < len(lines):
      m = re.search(r"Compiling entry function '([^']+)'", lines[i])
      if m:
        frame = used = ''
        for j in range(i + 1, min(i + 6, len(lines))):
          if 'bytes stack frame' in lines[j]:
            frame = lines[j].split(':', 1)[-1].strip() if 'ptxas info' in lines[j] else lines[j].strip()
          if 'Used ' in lines[j]:
            used = lines[j][lines[j].index('Used '):].strip()
            break
        rows.append((unit, m.group(1), used, frame))
      i += 1
  names = demangle([r[1] for r in rows])
  with open(os.path.join(ROOT, 'profiles', 'ptxas_summary.txt'), 'w') as f:
    f.write('# nvcc 12.9 -O3 -gencode arch=compute_100a,code=sm_100a -Xptxas -v: registers / barriers / static smem / '
            'spills per kernel\n# regenerate: python tools/ptxas_summary.py\n')
    for (unit, _, used, frame), n in sorted(zip(rows, names), key=lambda t: (t[0][0], t[1])):
      f.write('%-12s %-44s %s | %s\n' % (unit, n, used, frame))
  return len(rows)


MNEMONICS = [
    ('UTCHMMA', 'tcgen05.mma (kind::f16 / kind::tf32), one issuing thread'),
    ('UTCBAR', 'tcgen05.commit -> mbarrier'),
    ('LDTM', 'tcgen05.ld (TMEM -> registers)'),
    ('UTCATOMSWS', 'tcgen05.alloc / dealloc (TMEM columns)'),
    ('UTMALDG', 'cp.async.bulk.tensor (TMA tile load)'),
    ('UBLKCP', 'cp.async.bulk (1-D bulk copy)'),
    ('SYNCS', 'mbarrier arrive / try_wait'),
    ('LDGSTS', 'cp.async (16-byte global -> shared)'),
    ('UCGABAR', 'cluster barrier (barrier.cluster)'),
    ('ACQBULK', 'griddepcontrol.wait (PDL)'),
    ('PREEXIT', 'griddepcontrol.launch_dependents (PDL)'),
    ('FENCE.VIEW.ASYNC', 'fence.proxy.async'),
    ('REDG', 'red.global'),
    ('HMMA', 'legacy mma.sync (expected: 0)'),
]


def sass():
  txt = subprocess.run(['cuobjdump', '-sass', SO], capture_output=True, text=True).stdout
  per_fn = {}
  fn = None
  for line in txt.split('\n'):
    m = re.search(r'Function : (\S+)', line)
    if m:
      fn = m.group(1)
      continue
    if fn is None:
      continue
    for mn, _ in MNEMONICS:
      if re.search(r'\b' + re.escape(mn) + r'\b', line) or (mn + '.') in line:
        per_fn.setdefault(fn, {}).setdefault(mn, 0)
        per_fn[fn][mn] += 1
  fns = sorted(per_fn)
  names = dict(zip(fns, demangle(fns)))
  with open(os.path.join(ROOT, 'profiles', 'sass_evidence.txt'), 'w') as f:
    f.write('# cuobjdump -sass rec-attend-public_b200/librecattend_b200.so (sm_100a): instruction counts per kernel\n'
            '# regenerate: python tools/ptxas_summary.py\n#\n')
    for mn, what in MNEMONICS:
      tot = sum(v.get(mn, 0) for v in per_fn.values())
      f.write('# %-18s %-52s total %d\n' % (mn, what, tot))
    f.write('#\n')
    for fn in fns:
      f.write('%-60s %s\n' % (names[fn][:60], '  '.join('%s=%d' % kv for kv in sorted(per_fn[fn].items()))))
  return len(fns)


if __name__ == '__main__':
  print('ptxas entries:', ptxas())
  print('sass kernels with evidence mnemonics:', sass())
