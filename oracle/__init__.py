"""CPU oracle for the CryoVIT hot path -- TEST INFRASTRUCTURE ONLY.

A plain PyTorch fp32 restatement of the reference algorithm (VivianDLi/CryoVIT, /root/reference/src/cryovit)
for: slice pre-processing, the DINOv2-with-registers forward, the feature-volume layout/cast, the CryoVIT 3-D
head, and the masked Dice/F1 loss and metrics. Nothing under ``cryovit_b200/`` or ``cryovit/`` imports this
package: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs
may, and only as the checker or the reported CPU baseline -- never as the thing shipped.

Pinning status (see oracle/make_golden.py and DESIGN.md "Oracle"):
  * pre-processing, feature layout, head, loss/metrics, crop and collate: PINNED against the reference's own
    source files, executed in the build container with shims for the third-party packages that are not
    installed there (pytorch_lightning, torchmetrics, tensordict, h5py, hydra/omegaconf); the resulting golden
    vectors are committed under tests/golden/.
  * DINOv2 forward: the arithmetic lives in the un-vendored, un-pinned third-party ``facebookresearch/dinov2``
    (torch.hub default branch; entry ``dinov2_vitg14_reg``; reference call sites run/dino_features.py:25-28,58,336)
    and the reference holds no test or golden vector for it, so strictly PARITY IS UNPINNED for this part.
    The restatement in oracle/dinov2.py follows the published upstream algorithm and is cross-checked against
    the independent implementation transformers.Dinov2WithRegistersModel (same weights -> fp32 round-off
    agreement, tests/test_oracle.py); golden vectors from that cross-check are committed.
"""
