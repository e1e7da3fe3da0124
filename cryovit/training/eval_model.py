"""``python -m cryovit.training.eval_model model=cryovit datamodule=multi label_key=mito ...`` -- the reference's
entry point (training/eval_model.py:17-44): compose ``configs/eval_model.yaml``, validate, evaluate."""
import logging
import sys
import traceback
import warnings

from cryovit.config import compose, validate_experiment_config
from cryovit.run import eval_model

warnings.simplefilter("ignore")


def main(argv: list[str] | None = None) -> None:
    logging.basicConfig(level=logging.INFO, format="%(levelname)s %(message)s")
    cfg = compose("eval_model", list(sys.argv[1:] if argv is None else argv))
    validate_experiment_config(cfg, "eval_model")
    try:
        eval_model.run_trainer(cfg)
    except BaseException as err:  # noqa: BLE001
        logging.error("%s: %s", type(err).__name__, err)
        logging.error(traceback.format_exc())


if __name__ == "__main__":
    main()
