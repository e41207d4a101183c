"""RP_LOGGER set-up (reference utility/logger.py; not on the hot path)."""
import logging
import os
import sys
from datetime import datetime


def initialize_logger(config) -> logging.Logger:
    logger = logging.getLogger("RP_LOGGER")
    logger.handlers.clear()
    level = getattr(logging, str(config.debug.logging_level).upper(), logging.INFO)
    logger.setLevel(level)
    fmt = logging.Formatter("%(asctime)s\t%(filename)s\t\t%(funcName)s@%(lineno)d\t%(levelname)s\t%(message)s")
    stream = logging.StreamHandler(sys.stdout)
    stream.setLevel(level)
    stream.setFormatter(logging.Formatter("%(levelname)-8s [%(filename)s]: %(message)s"))
    logger.addHandler(stream)
    if config.debug.save_config:
        os.makedirs(config.general.path_logs, exist_ok=True)
        name = "%s_%s.log" % (config.general.name_scenario or "planner", datetime.now().strftime("%Y%m%d_%H%M%S"))
        fh = logging.FileHandler(os.path.join(config.general.path_logs, name))
        fh.setLevel(level)
        fh.setFormatter(fmt)
        logger.addHandler(fh)
    return logger
