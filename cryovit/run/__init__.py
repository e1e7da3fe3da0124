from . import dino_features, infer_model  # noqa: F401
