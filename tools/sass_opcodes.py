"""Per-kernel SASS opcode census of the product library: proof that the tcgen05 / TMEM / TMA paths are in the binary.

    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt

For every kernel in cryovit_b200/lib/libcryovit_b200.so (cuobjdump -sass): number of instructions and the counts of
UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UBLKCP (TMA loads), UTCBAR (tcgen05.commit), SYNCS
(mbarrier), UTCCP, HMMA (warp-level mma.sync), MUFU, LDGSTS (cp.async).
"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
so = ROOT / "cryovit_b200" / "lib" / "libcryovit_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
CLASSES = OrderedDict([("UTC*MMA", r"^UTC\w*MMA"), ("LDTM", r"^LDTM"), ("STTM", r"^STTM"), ("UTMALDG", r"^UTMALDG"),
                       ("UBLKCP", r"^UBLKCP"), ("UTCBAR", r"^UTCBAR"), ("SYNCS", r"^SYNCS"), ("UTCCP", r"^UTCCP"),
                       ("HMMA", r"^HMMA"), ("MUFU", r"^MUFU"), ("LDGSTS", r"^LDGSTS")])
kernels, cur = OrderedDict(), None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), Counter())
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["instructions"] += 1
        for name, pat in CLASSES.items():
            if re.match(pat, op):
                cur[name] += 1
print(f"# cuobjdump -sass {so.relative_to(ROOT)}: {len(kernels)} kernels (sm_100a). Columns: instructions, then opcode counts.")
print("# " + " ".join(["instr"] + list(CLASSES)))
for k, c in sorted(kernels.items(), key=lambda kv: -kv[1]["UTC*MMA"]):
    name = demangle(k)
    name = re.sub(r"\(.*", "", name)
    print(f"{name}\n    {c['instructions']:6d} " + " ".join(f"{n}={c[n]}" for n in CLASSES if c[n]))
tot = Counter()
for c in kernels.values():
    tot.update(c)
print("# total: " + " ".join(f"{n}={tot[n]}" for n in ["instructions"] + list(CLASSES)))
