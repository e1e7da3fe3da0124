from cryovit_b200.host.metrics import DiceLoss  # noqa: F401
