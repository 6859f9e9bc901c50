"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per kernel count, total, mean, share.
usage: python scripts/launch_summary.py gpurun_out/launches.csv [other.csv to compare]"""
import csv, re, sys, collections
def load(path):
    rows = []
    with open(path, newline="") as fh:
        lines = [l for l in fh if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum": continue
        v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
        us = v / 1000.0 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000.0)
        name = r["Kernel Name"]
        rows.append((name, us))
    return rows
def short(n):
    n = re.sub(r"\(.*$", "", n)                      # drop the argument list
    n = n.replace("mfm::", "").replace("(bool)", "")
    return n[:110]
def summarise(rows):
    agg = collections.OrderedDict()
    for n, us in rows:
        a = agg.setdefault(short(n), [0, 0.0]); a[0] += 1; a[1] += us
    return agg
a = summarise(load(sys.argv[1]))
b = summarise(load(sys.argv[2])) if len(sys.argv) > 2 else None
tot = sum(v[1] for v in a.values())
print(f"# {sys.argv[1]}: {sum(v[0] for v in a.values())} launches, {tot / 1000:.2f} ms of kernel time (serialised, cold-cache ncu replay: shares, not absolutes)")
print("| kernel | launches | total ms | mean us | share |" + (" other: launches | mean us |" if b else ""))
print("|---|---:|---:|---:|---:|" + ("---:|---:|" if b else ""))
for k, (c, t) in sorted(a.items(), key=lambda kv: -kv[1][1]):
    extra = ""
    if b is not None:
        o = b.get(k); extra = f" {o[0]} | {o[1] / o[0]:.1f} |" if o else " - | - |"
    print(f"| `{k}` | {c} | {t / 1000:.3f} | {t / c:.1f} | {100 * t / tot:.1f} % |{extra}")
