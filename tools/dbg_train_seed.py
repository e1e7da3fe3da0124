import sys, tempfile
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cryovit_b200.train import CryoVITHeadTrainerB200
from cryovit_b200.head import CryoVITHeadB200
from cryovit_b200.host import hdf
from cryovit_b200.host.fit import fit_head
from cryovit.datasets import TomoDataset
rng = np.random.default_rng(4)
tmp = Path(tempfile.mkdtemp())
recs, data = [], []
for i in range(9):
    feats = (0.1 * rng.standard_normal((384, 4, 3, 3))).astype(np.float16)
    feats[0] = 2.0 * np.sign(rng.standard_normal((4, 3, 3)))
    lab = (np.repeat(np.repeat(feats[0].astype(np.float32), 16, axis=1), 16, axis=2) > 0).astype(np.int8)
    lab[0, :8] = -1
    hdf.write_tomogram(tmp / "S" / f"t{i}.hdf", {"data": np.zeros((4, 48, 48), np.uint8), "labels/mito": lab, "dino_features": feats})
    recs.append({"sample": "S", "tomo_name": f"t{i}.hdf"})
    data.append((torch.from_numpy(feats).cuda(), torch.from_numpy(lab.astype(np.float32)).cuda()))
def held_out(sd):
    head = CryoVITHeadB200(384).load_state_dict(sd).cuda()
    ds = []
    for f, l in data[6:]:
        _, probs = head.segment_volume(f, want_logits=False)
        m = l > -1
        p = (probs >= 0.5).float()[m]
        ds.append(round(float(2 * (p * l[m]).sum() / (p.sum() + l[m].sum() + 1e-3)), 3))
    return ds
# (1) trainer directly, 6 tomograms x 30 epochs
tr = CryoVITHeadTrainerB200(384, lr=1e-3, seed=42)
losses = [float(tr.train_step(*data[i % 6])) for i in range(180)]
print("direct   : loss", [round(losses[k], 3) for k in (0, 30, 60, 120, 179)], "held-out dice", held_out(tr.state_dict()), flush=True)
# (2) through fit_head + TomoDataset(train=True)
ds = TomoDataset(recs[:6], "dino_features", "mito", "split_id", tmp, train=True)
for swa in (None, 24):
    sd = fit_head(ds, in_channels=384, max_epochs=30, lr=1e-3, swa_epoch_start=swa, seed=42)
    print(f"fit_head swa={swa}: held-out dice", held_out(sd), flush=True)
it = ds[0]
print("dataset item:", tuple(it.data.shape), it.data.dtype, tuple(it.label.shape), it.label.dtype, float(it.label.min()), float(it.label.max()))
