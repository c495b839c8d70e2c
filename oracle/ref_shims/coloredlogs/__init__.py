"""Import stub of coloredlogs (absent; model_training/utils.py:9-40)."""
import logging

DEFAULT_FIELD_STYLES = {}


def install(level=None, logger=None, field_styles=None, fmt=None, **kw):
    if logger is not None and level is not None:
        logger.setLevel(level if not isinstance(level, str) else getattr(logging, level))
