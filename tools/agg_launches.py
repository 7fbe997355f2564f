import csv, collections, sys
lines=[l for l in open(sys.argv[1]) if not l.startswith('==')]
rows=list(csv.DictReader(lines))
half=rows[len(rows)//2:]
agg=collections.defaultdict(lambda:[0,0.0])
for x in half:
    name=x['Kernel Name'].split('(')[0][:64]
    v=float(x['Metric Value'].replace(',','')); u=x['Metric Unit']
    v = v/1000 if u=='ns' else (v*1000 if u=='ms' else v)
    agg[name][0]+=1; agg[name][1]+=v
tot=sum(v[1] for v in agg.values())
print('launches in the second step: %d, serialized cold-cache GPU time %.0f us' % (len(half), tot))
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:int(sys.argv[2]) if len(sys.argv)>2 else 22]:
    print('%-66s n=%4d %9.1f us %5.1f%% avg %7.1f'%(k,v[0],v[1],100*v[1]/tot, v[1]/v[0]))
