"""``python -m cryovit.training.dino_features [key=value ...]`` -- the reference's entry point
(training/dino_features.py:16-41): compose ``configs/dino_features.yaml`` with the command-line overrides, validate,
run; any error of the run is logged with its traceback and swallowed, as the reference does (:33-37)."""
import logging
import sys
import traceback
import warnings

from cryovit.config import compose, validate_dino_config
from cryovit.run import dino_features

warnings.simplefilter("ignore")


def main(argv: list[str] | None = None) -> None:
    logging.basicConfig(level=logging.DEBUG, format="%(levelname)s %(message)s")
    cfg = compose("dino_features", list(sys.argv[1:] if argv is None else argv))
    validate_dino_config(cfg)
    try:
        dino_features.run_trainer(cfg)
    except BaseException as err:  # noqa: BLE001
        logging.error("%s: %s", type(err).__name__, err)
        logging.error(traceback.format_exc())


if __name__ == "__main__":
    main()
