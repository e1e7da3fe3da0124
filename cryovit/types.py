from cryovit_b200.host.datasets import BatchedTomogramData, BatchedTomogramMetadata, TomogramData  # noqa: F401
from cryovit_b200.host.model_io import ModelType  # noqa: F401,E402
