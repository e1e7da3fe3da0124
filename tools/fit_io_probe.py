"""The head's fit loop on tomogram FILES at BASELINE size: seconds per epoch with the training set resident in HBM
(host/fit.py default) against re-reading every file in every epoch (cache_gb=0, what the reference's loader does).
usage: python tools/fit_io_probe.py [n_tomograms=6] [epochs=4]"""
import logging
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from cryovit_b200.host import fit, hdf  # noqa: E402
from cryovit_b200.host.datasets import TomoDataset  # noqa: E402

n, epochs = (int(sys.argv[1]) if len(sys.argv) > 1 else 6), (int(sys.argv[2]) if len(sys.argv) > 2 else 4)
logging.basicConfig(level=logging.INFO, format="%(message)s")
root = Path("/tmp/cryovit_fit_probe")
rng = np.random.default_rng(0)
recs = []
for i in range(n):
    p = root / "S" / f"t{i}.hdf"
    if not p.exists():
        feats = rng.standard_normal((1536, 128, 32, 32), dtype=np.float32).astype(np.float16)
        lab = (rng.random((128, 512, 512)) > 0.7).astype(np.int8)
        hdf.write_tomogram(p, {"data": np.zeros((128, 512, 512), np.uint8), "labels/mito": lab, "dino_features": feats})
    recs.append({"sample": "S", "tomo_name": f"t{i}.hdf"})
ds = TomoDataset(recs, "dino_features", "mito", "split_id", root, train=True)
for label, gb in (("resident in HBM", None), ("files every epoch", 0.0)):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fit.fit_head(ds, in_channels=1536, max_epochs=epochs, swa_epoch_start=None, cache_gb=gb)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"[fit-io] {label}: {epochs} epochs x {n} tomograms in {dt:.2f} s = {dt / (epochs * n) * 1e3:.1f} ms per step", flush=True)
