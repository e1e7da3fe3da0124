from cryovit_b200.host.metrics import DiceMetric, F1Metric  # noqa: F401
