"""Aggregate `ncu --page source --csv --print-source cuda,sass` output per CUDA source line:
samples, executed warp instructions and the dominant stall reasons.  Usage: ncu_source_lines.py report.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = collections.defaultdict(lambda: collections.Counter())
text = {}
fpath = None; hdr = None; cur = None
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fpath = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None: continue
    d = {}
    for k, v in zip(hdr, r):
        if k not in d: d[k] = v
    if d.get('Line No', '') != '':
        cur = (fpath, int(d['Line No'])); text[cur] = r[1].strip()[:110]
        continue
    if cur is None: continue
    a = agg[cur]
    def num(k):
        try: return float(d.get(k, '0') or 0)
        except ValueError: return 0.0
    a['samples'] += num('# Samples'); a['inst'] += num('Instructions Executed')
    for k in hdr:
        if k.startswith('stall_') and 'Not Issued' not in k: a[k] += num(k)
tot = sum(a['samples'] for a in agg.values()) or 1
print('total samples %d, warp instructions %d' % (tot, sum(a['inst'] for a in agg.values())))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]['samples'])[:top]:
    st = sorted(((k[6:], v) for k, v in a.items() if k.startswith('stall_')), key=lambda kv: -kv[1])[:3]
    print('%5.1f%%  inst %11d  %-18s:%4d  %-44s %s' % (100 * a['samples'] / tot, a['inst'], key[0], key[1],
          ' '.join('%s=%d' % (k, v) for k, v in st), text[key]))
