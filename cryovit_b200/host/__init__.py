"""Host-side mirror of the reference's interfaces for the hot path (Python, like the reference): config
composition, the tomogram file layout, datasets / collate, loss + metrics, the feature-extraction runner and the
slice / tomogram sharding across GPUs. The arithmetic is in the sm_100a library behind ``cryovit_b200.ops``."""
