"""Launch-list CSV of `ncu --metrics gpu__time_duration.sum --csv` -> markdown share table (profiles/*.md).

usage: python tools/summarize_launches.py gpurun_out/r01_launches_vN.csv > table.md"""
import collections
import csv
import sys


def summarize(path):
    rows = list(csv.reader(line for line in open(path) if line.startswith('"')))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    total = sum(a[1] for a in agg.values())
    out = ["| launches | total ms | share | kernel |", "|---:|---:|---:|---|"]
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        short = name.split("(")[0].replace("void ", "")
        out.append(f"| {n} | {ns / 1e6:.2f} | {ns / total:.1%} | `{short}` |")
    return "\n".join(out), len(rows) - 1, total / 1e6


if __name__ == "__main__":
    table, n, ms = summarize(sys.argv[1])
    print(f"{n} launches, {ms:.1f} ms under ncu\n\n{table}")
