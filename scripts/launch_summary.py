"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and the
kernel sequence of one FM update / one MALA iteration / the flow-MH iteration.
usage: python scripts/launch_summary.py gpurun_out/launches.csv [--phase]"""
import collections
import csv
import re
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = []
    for x in csv.DictReader(lines):
        rows.append((int(x["ID"]), x["Kernel Name"], x["Grid Size"], float(x["Metric Value"])))
    return rows


def short(name):
    name = re.sub(r"^void ", "", name)
    m = re.match(r"([\w:]+)<([^>]*?)mfm::(\w+)>", name)
    if m:
        flags = ",".join(re.findall(r"\b(\d)\b", m.group(2)))
        return f"{m.group(1).split('::')[-1]}<{flags},{m.group(3)}>"
    return re.sub(r"\(.*", "", name).split("::")[-1][:48]


def table(rows, title):
    agg = collections.OrderedDict()
    for _, k, _, t in rows:
        a = agg.setdefault(short(k), [0, 0.0]); a[0] += 1; a[1] += t
    tot = sum(v[1] for v in agg.values())
    print(f"\n{title}: {len(rows)} launches, {tot / 1e6:.2f} ms")
    print("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1] / 1e6:.2f} | {100 * v[1] / tot:.1f}% | {v[1] / v[0] / 1e3:.1f} |")


def main():
    rows = load(sys.argv[1])
    table(rows, "all launches")
    fm = [i for i, r in enumerate(rows) if "fm_batch" in r[1]]
    if len(fm) >= 2 and "--phase" in sys.argv:
        # last complete outer iteration = FM update followed by the data generator of the next one
        table(rows[fm[-2]:fm[-1]], "one FM update + the following data-generator call (last complete pair)")
        gaps = [fm[i + 1] - fm[i] for i in range(len(fm) - 1)]
        big = max(range(len(gaps)), key=lambda i: gaps[i])
        table(rows[fm[big]:fm[big + 1]], "the FM update followed by the flow-MH iteration")


if __name__ == "__main__":
    main()
