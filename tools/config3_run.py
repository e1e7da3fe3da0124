"""BASELINE config 3 for real: a generated sample directory of 128x512x512 uint8 tomogram files on the box's local disk,
`torchrun --nproc-per-node N -m cryovit.training.dino_features` over it (tomograms dealt round-robin to the ranks, reader /
writer threads around the GPU extractor, gzip of `data`, 403 MB of features per result file), wall clock per rank from the
runner's own log and for the whole command.  usage: python tools/config3_run.py [n_gpus=8] [n_tomograms=64]"""
import json
import os
import re
import shutil
import subprocess
import sys
import time
from concurrent.futures import ProcessPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
D, H, W = 128, 512, 512


def make(args):
    path, seed = args
    from cryovit_b200.host import hdf

    hdf.write_tomogram(path, {"data": np.random.default_rng(seed).integers(0, 256, (D, H, W), dtype=np.uint8)})
    return path.stat().st_size


def main():
    n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    n_tomo = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    base = Path(os.environ.get("CFG3_DIR", "/tmp/cryovit_cfg3"))
    shutil.rmtree(base, ignore_errors=True)
    free = shutil.disk_usage(base.parent).free
    need = n_tomo * (34e6 + 440e6)
    while need > 0.8 * free and n_tomo > n_gpus:
        n_tomo //= 2
        need = n_tomo * (34e6 + 440e6)
    src = base / "dino_features" / "Q18"
    src.mkdir(parents=True)
    t0 = time.perf_counter()
    with ProcessPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
        sizes = list(ex.map(make, [(src / f"tomo_{i:03d}.hdf", 1000 + i) for i in range(n_tomo)]))
    t_gen = time.perf_counter() - t0
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n_gpus}", "--master-addr", "127.0.0.1",
           "--master-port", "29711", "-m", "cryovit.training.dino_features", f"paths.data_dir={base}", f"paths.exp_dir={base}/exp",
           f"paths.model_dir={base}/models", "sample=Q18", "batch_size=128", "+allow_random_weights=true"]
    env = dict(os.environ, PYTHONPATH=str(ROOT) + os.pathsep + os.environ.get("PYTHONPATH", ""))
    t0 = time.perf_counter()
    r = subprocess.run(cmd, cwd=str(ROOT), env=env, capture_output=True, text=True)
    wall = time.perf_counter() - t0
    log = r.stdout + r.stderr
    per_rank = [(int(m.group(1)), int(m.group(2)), float(m.group(3)), m.group(4))
                for m in re.finditer(r"rank (\d+)/\d+: (\d+) tomograms of Q18 in ([0-9.]+) s \(files in -> files out, ([\w-]+) container\)", log)]
    out_files = sorted((base / "tomograms" / "Q18").glob("*.hdf"))
    res = {
        "workload": f"BASELINE config 3: {n_tomo} x ({D},{H},{W}) uint8 tomogram files -> ViT-g/14 features, {n_gpus} GPUs, "
                    "python -m cryovit.training.dino_features under torchrun (random-init weights, +allow_random_weights=true)",
        "n_gpus": n_gpus, "tomograms": n_tomo, "slices": n_tomo * D, "returncode": r.returncode,
        "files_written": len(out_files), "bytes_written": sum(f.stat().st_size for f in out_files),
        "source_bytes": int(sum(sizes)), "source_generation_s": round(t_gen, 1),
        "wall_s_whole_command": round(wall, 2), "slices_per_s_whole_command": round(n_tomo * D / wall, 1),
        "per_rank": [{"rank": a, "tomograms": b, "seconds": c, "container": d} for a, b, c, d in sorted(per_rank)],
        "disk": str(base.parent), "disk_free_GB_before": round(free / 1e9, 1),
    }
    if per_rank:
        slowest = max(c for _, _, c, _ in per_rank)
        res["seconds_files_in_to_files_out"] = slowest  # max over ranks, model construction excluded
        res["slices_per_s_files_in_to_files_out"] = round(n_tomo * D / slowest, 1)
    else:
        res["log_tail"] = log[-3000:]
    print(json.dumps(res, indent=1))
    Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / f"config3_n{n_gpus}.json").write_text(json.dumps(res, indent=1))
    shutil.rmtree(base, ignore_errors=True)


if __name__ == "__main__":
    main()
