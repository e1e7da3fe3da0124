from cryovit_b200.host.config import (  # noqa: F401
    DINO_PATCH_SIZE, MISSING, Cfg, compose, instantiate, missing_keys, samples, tomogram_exts, validate_dino_config, validate_experiment_config,
)
