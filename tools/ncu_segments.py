"""Per-segment view of `ncu --page source --csv`: runs of SASS with the same execution count, their share of the
issued warp instructions and the average active lanes (where the lanes are lost)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
ie, it, isamp, isrc = (hdr.index(k) for k in ('Instructions Executed', 'Thread Instructions Executed', '# Samples', 'Source'))
tot = sum(int(r[ie]) for r in data)
tsamp = sum(int(r[isamp]) for r in data)
print(f'warp instructions {tot}  lanes/inst {sum(int(r[it]) for r in data) / tot:.2f}  SASS lines {len(data)}  samples {tsamp}')
segs, cur = [], None
for k, r in enumerate(data):
    e, t, s = int(r[ie]), int(r[it]), int(r[isamp])
    if cur and abs(e - cur['e']) <= 0.02 * max(e, cur['e']):
        cur['n'] += 1; cur['E'] += e; cur['T'] += t; cur['S'] += s; cur['end'] = k
    else:
        if cur:
            segs.append(cur)
        cur = {'start': k, 'end': k, 'e': e, 'n': 1, 'E': e, 'T': t, 'S': s}
segs.append(cur)
for s in segs:
    if s['E'] > 0.005 * tot:
        print(f"sass[{s['start']:5d}-{s['end']:5d}] n={s['n']:4d} exec/inst={s['e']:9d} inst share={100 * s['E'] / tot:5.1f}% "
              f"lanes={s['T'] / s['E']:5.1f} stall-sample share={100 * s['S'] / tsamp:5.1f}%")
