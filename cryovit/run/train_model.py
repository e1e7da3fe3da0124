from cryovit_b200.host.train_model import build_datamodule, run_trainer, setup_exp_dir  # noqa: F401
