"""Drop-in module paths of VivianDLi/CryoVIT for the B200 hot path: ``cryovit.training.dino_features``,
``cryovit.run.dino_features``, ``cryovit.datasets.VITDataset``, ``cryovit.models.CryoVIT``, ``cryovit.config``.
Everything here re-exports ``cryovit_b200.host`` (host logic) over the sm_100a library; nothing else of the
reference package is provided."""
