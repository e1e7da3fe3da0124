from cryovit_b200.host.datasets import TomoDataset, VITDataset  # noqa: F401
