from cryovit_b200.host.datasets import BatchedTomogramData, BatchedTomogramMetadata, TomogramData  # noqa: F401
