"""cryovit_b200: B200-native (sm_100a) implementation of CryoVIT's feature-extraction + 3-D head hot path."""

__version__ = "0.1.0"
