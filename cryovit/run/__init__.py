from . import dino_features  # noqa: F401
