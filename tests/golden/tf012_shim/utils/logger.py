"""Stand-in for the reference's utils/logger.py (only `get()` is used at import time by modellib.py)."""
import logging


def get(*args, **kwargs):
  return logging.getLogger('rec-attend-reference')
