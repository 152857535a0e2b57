"""Logger of the package (same name as /root/reference/polars_bio/logging.py:10-41)."""
import logging

logger = logging.getLogger("polars_bio")


def set_loglevel(level: str) -> None:
    level = level.lower()
    levels = {"debug": logging.DEBUG, "info": logging.INFO, "warn": logging.WARNING, "warning": logging.WARNING}
    if level not in levels:
        raise ValueError(f"Invalid log level: {level}")
    logger.setLevel(levels[level])
