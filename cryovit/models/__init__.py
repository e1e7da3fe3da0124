from cryovit_b200.host.metrics import DiceLoss, DiceMetric, F1Metric  # noqa: F401
from cryovit_b200.host.models import CryoVIT  # noqa: F401
