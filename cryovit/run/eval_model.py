from cryovit_b200.host.eval_model import run_trainer, test_step  # noqa: F401
from cryovit_b200.host.train_model import setup_exp_dir  # noqa: F401
