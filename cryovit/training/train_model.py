"""``python -m cryovit.training.train_model model=cryovit datamodule=multi label_key=mito datamodule.sample=[A,B] ...``
-- the reference's entry point (training/train_model.py:17-55): compose ``configs/train_model.yaml``, validate the
mandatory keys, train; an error of the run is logged with its traceback and swallowed, as the reference does."""
import logging
import sys
import traceback
import warnings

from cryovit.config import compose, validate_experiment_config
from cryovit.run import train_model

warnings.simplefilter("ignore")


def main(argv: list[str] | None = None) -> None:
    logging.basicConfig(level=logging.INFO, format="%(levelname)s %(message)s")
    cfg = compose("train_model", list(sys.argv[1:] if argv is None else argv))
    validate_experiment_config(cfg, "train_model")
    try:
        train_model.run_trainer(cfg)
    except BaseException as err:  # noqa: BLE001
        logging.error("%s: %s", type(err).__name__, err)
        logging.error(traceback.format_exc())


if __name__ == "__main__":
    main()
