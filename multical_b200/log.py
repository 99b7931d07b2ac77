"""Logger named like the reference's (multical/io/logging.py:11-23) so that existing handlers
(`MemoryHandler`, workspace log files) keep receiving the solver's iteration table."""
import logging

logger = logging.getLogger("calibration")


def info(msg, *args, **kwargs): return logger.info(msg, *args, **kwargs)
def debug(msg, *args, **kwargs): return logger.debug(msg, *args, **kwargs)
def warning(msg, *args, **kwargs): return logger.warning(msg, *args, **kwargs)
