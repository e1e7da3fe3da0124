"""Stall-reason summary of one kernel from `ncu -i X.ncu-rep --page source --csv` (SASS view with sampling columns).
usage: python tools/ncu_source_stalls.py X.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def f(r, k):
    try:
        return float(r[ix[k]].replace(",", ""))
    except (ValueError, IndexError):
        return 0.0


tot = sum(f(r, "# Samples") for r in data)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {c: sum(f(r, c) for r in data) for c in stall_cols}
print(rows[0][1][:150])
print("total samples", tot, "| warp instructions executed", sum(f(r, "Instructions Executed") for r in data))
for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {c:28s} {v:9.0f} {v / tot:.1%}")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
print(f"top {n} instructions by samples (samples, executed, SASS, two largest stall reasons)")
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:n]:
    st = sorted(((c, f(r, c)) for c in stall_cols), key=lambda kv: -kv[1])[:2]
    print(f"{f(r, '# Samples'):7.0f} {f(r, 'Instructions Executed'):10.0f}  {r[ix['Source']][:80]:80s} {st[0][0]}={st[0][1]:.0f} {st[1][0]}={st[1][1]:.0f}")
# by opcode
ops = {}
for r in data:
    src = r[ix["Source"]].split()
    op = next((t for t in src if not t.startswith("@")), "?").split(".")[0]
    a = ops.setdefault(op, [0.0, 0.0])
    a[0] += f(r, "# Samples")
    a[1] += f(r, "Instructions Executed")
print("by opcode (samples share, executed share)")
ex = sum(a[1] for a in ops.values())
for op, a in sorted(ops.items(), key=lambda kv: -kv[1][0])[:18]:
    print(f"  {op:12s} {a[0] / tot:6.1%} {a[1] / ex:6.1%}")
