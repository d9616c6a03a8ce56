"""Re-base the stamps of a profiles/gf_phase_clock.py output to the first stamp: python tools/pc_rebase.py FILE"""
import re
import sys
rows = [l for l in open(sys.argv[1]) if 'tile' in l and 'whole' not in l]
vals = [int(x) for l in rows for x in re.findall(r'\d{10,}', l)]
vals = [v for v in vals if v > max(vals) - 10 ** 7]
t0 = min(vals)
for l in rows:
    f = l.split()
    if f[2] == 'top':
        print('top', int(f[3]) - t0)
        continue
    a, b, e = int(f[3]) - t0, int(f[4]) - t0, int(f[5]) - t0
    print(f[0], f[1], f"{f[2]:5s}", f"{a:7d} {b:7d} {e:7d}   wait {b - a:6d}  work {e - b:6d}",
          ('ring-wait ' + l.split('ring-wait')[1].split()[0]) if 'ring-wait' in l else '')
