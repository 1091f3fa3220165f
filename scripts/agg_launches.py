import csv, collections, sys
for f in sys.argv[1:]:
    with open(f) as fh:
        lines=[l for l in fh if not l.startswith('==')]
    agg=collections.OrderedDict()
    for row in csv.DictReader(lines):
        k=row['Kernel Name'][:70]; v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
        v = v/1e3 if u=='ns' else (v*1e3 if u=='ms' else v)
        a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
    tot=sum(a[1] for a in agg.values())
    print(f, 'total us %.1f' % tot)
    for k,(n,t) in sorted(agg.items(), key=lambda x:-x[1][1]):
        print(f"  {k:70s} n={n:4d} total={t:10.1f}us avg={t/n:9.2f}us share={t/tot:.3f}")
