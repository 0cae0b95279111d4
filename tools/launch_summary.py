import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
cols = rows[hdr]; data = rows[hdr + 1:]
ki, vi, ui = cols.index('Kernel Name'), cols.index('Metric Value'), cols.index('Metric Unit')
agg = collections.OrderedDict()
for r in data:
    if len(r) <= vi: continue
    name = re.sub(r'\(.*', '', r[ki]); name = re.sub(r'void |\(anonymous namespace\)::|<unnamed>::', '', name)
    v = float(r[vi].replace(',', '')); u = r[ui]
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"total {tot/1e3:.3f} ms over {sum(a[0] for a in agg.values())} launches")
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t/1e3:9.3f} ms {100*t/tot:5.1f}%  n={c:5d}  avg {t/c:9.1f} us  {n[:90]}")
