"""Test-only shim for `logzero.logger` (inference.py:11)."""
import logging
logger = logging.getLogger("logzero_shim")
