import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; acc = collections.defaultdict(list)
for r in rows:
    if hdr is None:
        if "Kernel Name" in r: hdr = r
        continue
    if len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") == "gpu__time_duration.sum":
        v = float(d["Metric Value"].replace(",", "")); u = d["Metric Unit"]
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1.0)
        acc[d["Kernel Name"][:60]].append(v)
for k, v in sorted(acc.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:60s} n={len(v):4d} total {sum(v):9.3f} ms avg {sum(v)/len(v):8.4f} ms")
