from . import dino_features, eval_model, infer_model, train_model  # noqa: F401
