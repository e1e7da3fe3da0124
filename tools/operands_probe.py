"""Error and throughput of ViT-g/14-reg4 operand formats on one box, interleaved: per-token relative error of 4 slices
(448 x 448, weight seeds 0 and 11, one heavy-tailed x10 set) against the fp32 oracle evaluated on the GPU, and slices/s of
the device-resident 128 x 512 x 512 extraction (the bench's `value` leg).  usage: python tools/operands_probe.py mixed mixed-attn ..."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from cryovit_b200.extract import extract_tomogram_device  # noqa: E402
from cryovit_b200.vit import CONFIGS, DinoVisionTransformerB200, random_state_dict  # noqa: E402
from test_gpu_parity import _oracle_fp32_on_gpu, _percentiles, heavy_tailed_state_dict  # noqa: E402

modes = sys.argv[1:] or ["mixed", "mixed-attn"]
cfg = CONFIGS["dinov2_vitg14_reg"]
x = torch.rand(4, 3, 448, 448, generator=torch.Generator().manual_seed(1))
cases = [("seed 0", random_state_dict(cfg, seed=0)), ("seed 11", random_state_dict(cfg, seed=11)),
         ("heavy-tailed x10", heavy_tailed_state_dict(cfg, 21, 10.0))]
refs = []
for name, sd in cases:
    sdg = {k: v.cuda() for k, v in sd.items()}
    refs.append(_oracle_fp32_on_gpu(sdg, x, cfg.num_heads))
    del sdg
    torch.cuda.empty_cache()
for mode in modes:
    for (name, sd), ref in zip(cases, refs):
        model = DinoVisionTransformerB200(cfg, mode).load_state_dict(sd).cuda()
        got = model.forward_features(x.cuda())["x_norm_patchtokens"].float().cpu()
        rmax, rmean, r999, cmin = _percentiles(got, ref)
        print(f"[{mode}] {name}: rel-err max {rmax:.3e} mean {rmean:.3e} 99.9th {r999:.3e} min cosine {cmin:.6f}", flush=True)
        del model
        torch.cuda.empty_cache()
tomo = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (128, 512, 512), dtype=np.uint8)).cuda()
sd0 = cases[0][1]
for rnd in range(2):
    for mode in modes:
        model = DinoVisionTransformerB200(cfg, mode).load_state_dict(sd0).cuda()
        feats = None
        for _ in range(3):
            feats = extract_tomogram_device(tomo, model, 128, out=feats)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 6
        for _ in range(n):
            feats = extract_tomogram_device(tomo, model, 128, out=feats)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
        print(f"[{mode}] round {rnd}: {128 / dt:.1f} slices/s ({dt * 1e3:.1f} ms per tomogram)", flush=True)
        del model, feats
        torch.cuda.empty_cache()
