"""visco_b200 — B200 (sm_100a) drop-in for VISCO's SVD -> truncate -> reconstruct hot path.

Mirrors the names the reference exposes for this path (reference visco/__init__.py:8-32):
``__version__``, ``PCKGDIR``, ``get_logger``, ``LOG``; the hot-path callables live in
``visco_b200.compress_ms`` / ``visco_b200.decompress_ms`` under the reference's own names.
"""
import logging
import os

__version__ = "0.1.0"
PCKGDIR = os.path.dirname(os.path.abspath(__file__))


def get_logger(name, level="INFO"):
    """Same logger name/format as the reference (visco/__init__.py:15-28) without touching the root config."""
    if isinstance(level, str):
        level = getattr(logging, level, logging.INFO)
    log = logging.getLogger(name)
    if not log.handlers:
        h = logging.StreamHandler()
        h.setFormatter(logging.Formatter("%(asctime)s-%(name)s-%(levelname)-8s| %(message)s", datefmt="%m:%d %H:%M:%S"))
        log.addHandler(h)
        log.propagate = False
    log.setLevel(level)
    return log


LOG = get_logger("VISCO")
