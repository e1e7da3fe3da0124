from cryovit_b200.host.model_io import ModelType, SavedModel, load_model, save_model  # noqa: F401
