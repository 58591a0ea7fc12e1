import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = collections.defaultdict(lambda: [0.0,0])
for r in rows:
    if hdr is None:
        if 'Kernel Name' in r: hdr = r
        continue
    d = dict(zip(hdr, r))
    if d.get('Metric Name') != 'gpu__time_duration.sum': continue
    name = re.sub(r'\(.*', '', d['Kernel Name']).replace('(anonymous namespace)::','').replace('void <unnamed>::','')
    v = float(d['Metric Value'].replace(',','')); unit = d['Metric Unit']
    if unit == 'ns': v/=1e3
    elif unit == 'ms': v*=1e3
    key=(name, d['Grid Size'], d['Block Size'])
    agg[key][0]+=v; agg[key][1]+=1
tot=sum(v[0] for v in agg.values())
print("total us %.1f kernels %d" % (tot, sum(v[1] for v in agg.values())))
for k,(us,n) in sorted(agg.items(), key=lambda kv:-kv[1][0])[:int(sys.argv[2]) if len(sys.argv)>2 else 30]:
    print(f"{us:10.1f} us {100*us/tot:5.1f}%  n={n:4d} avg={us/n:8.2f}  {k}")
