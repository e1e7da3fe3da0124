from cryovit_b200.host.callbacks import BatchedModelResult, CsvWriter, TestPredictionWriter  # noqa: F401
