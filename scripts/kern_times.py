import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=None; agg=collections.defaultdict(lambda: collections.defaultdict(float)); cnt=collections.Counter()
for r in rows:
    if "Kernel Name" in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); k=d["Kernel Name"][:56]
        try: agg[k][d["Metric Name"].split("__")[-1][:18]]+=float(d["Metric Value"].replace(",",""))
        except: pass
        if d["Metric Name"]=="gpu__time_duration.sum": cnt[k]+=1
for k,v in agg.items(): print(cnt[k], k, {a: round(b/1e6,3) if "time" in a else round(b/1e9,3) for a,b in v.items()})
