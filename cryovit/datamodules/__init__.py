from cryovit_b200.host.datamodules import (  # noqa: F401
    BaseDataModule, MultiSampleDataModule, SingleSampleDataModule, TomoLoader,
)
