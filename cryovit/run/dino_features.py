from cryovit_b200.host.dino_features import (  # noqa: F401
    _dino_features, _process_sample, _save_data, dino_model, load_model, run_trainer,
)
