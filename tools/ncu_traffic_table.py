"""Curated DRAM-traffic table for bench.py's roofline blocks: the `ncu --set full` extracts of one build
(profiles/<round>_ncu_full_<tag>_{hot,head,train}_kernels.json, launch order) matched with the C-ABI calls that made
them (the per-layer tables of a bench line of the same build: head.layers / head_train.kernels).

    python tools/ncu_traffic_table.py r02 v1 gpurun_out/bench_full.json > profiles/r02_ncu_traffic.json
"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
rnd, tag, bench = sys.argv[1], sys.argv[2], json.load(open(sys.argv[3]))
build = f"{rnd} {tag}"


def load(part):
    return json.load(open(ROOT / "profiles" / f"{rnd}_ncu_full_{tag}_{part}_kernels.json"))


def gb(r):
    return round(r["dram_traffic_GB"] * 1e9)


out = []
# ViT: tools/ncu_target.py launches layernorm, qkv, attention, proj, w12, w3 in this order
for label, r in zip(["layernorm", "qkv_gemm", "attention", "proj_gemm", "w12_swiglu", "w3_gemm"], load("hot")):
    out.append({"kernel": label, "dims": None, "dram_bytes": gb(r), "ncu_name": r["name"][:60], "build": build})
# head forward: one ncu record per kernel in call order; a groupnorm_fold call is two kernels (finalize + fold), a conv3d_rows
# call with 32 output channels two passes
recs = load("head")
i = 0
for row in bench["head"]["layers"]:
    n = 2 if row["kernel"] == "groupnorm_fold" or (row["kernel"] == "conv3d_rows_ndhwc" and row["dims"][4] == 32) else 1  # 2 passes of 16
    out.append({"kernel": row["kernel"], "dims": row["dims"], "dram_bytes": sum(gb(r) for r in recs[i:i + n]),
                "ncu_name": recs[i + n - 1]["name"][:60], "build": build})
    i += n
assert i == len(recs), (i, len(recs))
# training: the tcgen05 weight-gradient launches of one step in backward order
recs = load("train")
order = [("wgrad_tc8_ndhwc", None), ("wgrad_tc8_ndhwc", None), ("wgrad_mn_ndhwc", (1, 1, 2097152)), ("wgrad_tcn_ndhwc", (128, 256, 256, 16, 16)),
         ("wgrad_tcn_ndhwc", (128, 256, 256, 32, 16)), ("wgrad_mn_ndhwc", (1, 1, 1048576)), ("wgrad_tcn_ndhwc", (128, 128, 128, 32, 32, 4)),
         ("wgrad_tcn_ndhwc", (128, 128, 128, 32, 32, 8)), ("wgrad_mn_ndhwc", (1, 1, 524288)), ("wgrad_mn_ndhwc", (128, 64, 64, 64, 64)),
         ("wgrad_mn_ndhwc", (128, 64, 64, 128, 64)), ("wgrad_mn_ndhwc", (1, 1, 131072)), ("wgrad_mn_ndhwc", (128, 32, 32, 192, 192)),
         ("wgrad_mn_ndhwc", (128, 32, 32, 1024, 192))]
rows = bench["head_train"]["kernels"]
seen = set()
for (kern, prefix), r in zip(order, recs):
    match = [row for row in rows if row["kernel"] == kern and (prefix is None or tuple(row["dims"][:len(prefix)]) == prefix)]
    assert match, (kern, prefix)
    key = (kern, tuple(match[0]["dims"]))
    if key in seen:
        continue
    seen.add(key)
    out.append({"kernel": kern, "dims": match[0]["dims"], "dram_bytes": gb(r), "ncu_name": r["name"][:60], "build": build})
print(json.dumps(out, indent=1))
