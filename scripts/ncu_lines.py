"""Stall samples per CUDA source line of an .ncu-rep captured with --import-source on (read here with `ncu -i`)."""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows[:8]) if '# Samples' in r][0]
hdr = rows[hi]
k = hdr.index('# Samples')
first = hdr.index('stall_barrier')
names = hdr[first:first + 17]
data = []
for r in rows[hi + 1:]:
    if len(r) < k + 1 or r[0] == '':
        continue
    try:
        data.append((int(r[k]), r[0], r[1], [int(x) if x.isdigit() else 0 for x in r[first:first + 17]]))
    except ValueError:
        pass
tot = sum(d[0] for d in data)
print('total samples', tot)
for n, ln, src, st in sorted(data, key=lambda d: -d[0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 16]:
    print("%5.1f%% L%s: %s | %s" % (100 * n / tot, ln, src.strip()[:100], {nm[6:]: v for nm, v in zip(names, st) if v > n * 0.15}))
