"""Summarise an ncu --csv metrics log per kernel: count, total/avg duration, share, and any extra metrics averaged."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = None
per = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r:
        hdr = r; continue
    if not hdr or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("unnamed>::", "").replace("flk::", "")
    tmpl = d["Kernel Name"].split("(")[0]
    name = tmpl.split("::")[-1]
    key = (name, d["Grid Size"], d["Block Size"])
    try:
        v = float(d["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    per.setdefault(key, collections.defaultdict(list))[d["Metric Name"]].append(v)
tot = sum(sum(m.get("gpu__time_duration.sum", [0])) for m in per.values())
print("%-44s %-16s %5s %9s %6s  %s" % ("kernel", "grid", "n", "avg us", "share", "other metrics (avg)"))
for (name, grid, blk), m in sorted(per.items(), key=lambda kv: -sum(kv[1].get("gpu__time_duration.sum", [0]))):
    t = m.get("gpu__time_duration.sum", [0])
    unit = 1e3  # ns -> us
    others = "  ".join("%s=%.3g" % (k.replace("__", ".").split(".")[-2] + "." + k.split(".")[-1] if False else k, sum(v) / len(v)) for k, v in m.items() if k != "gpu__time_duration.sum")
    print("%-44s %-16s %5d %9.1f %5.1f%%  %s" % (name[:44], grid.replace(" ", ""), len(t), sum(t) / len(t) / unit, 100 * sum(t) / tot if tot else 0, others))
print("total %.1f us" % (tot / 1e3))
