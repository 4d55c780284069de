"""Opcode histogram of an `ncu --page source --csv` dump, weighted by executed warp-instructions."""
import csv, collections, re, sys
path, npx = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 48e6
rows = list(csv.reader(open(path)))
hdr = None
for i, r in enumerate(rows):
    if 'Source' in r:
        hdr, start = r, i + 1
        break
si = hdr.index('Source')
ei = [j for j, c in enumerate(hdr) if 'Instructions Executed' in c and 'Thread' not in c and 'Pred' not in c][0]
tot, n, first = collections.Counter(), 0, True
kern = 0
for r in rows[start:]:
    if len(r) <= max(si, ei):
        continue
    try:
        cnt = int(float(r[ei]))
    except ValueError:
        if 'Source' in r:
            kern += 1
        continue
    if kern > 0:
        break
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[si])
    if not m:
        continue
    tot[m.group(2)] += cnt
    n += cnt
print('total warp-inst', n, ' thread-inst per px %.1f' % (n * 32 / npx))
for op, c in tot.most_common(40):
    print('%-10s %12d  %5.1f%%  per-px %.1f' % (op, c, 100 * c / n, c * 32 / npx))
