"""hamt_b200 -- Blackwell-native (sm_100a) implementation of the HAMT data-parallel hot path.

The directory is named ``vln-hamt_b200`` (not importable as such); import it as ``hamt_b200``
through the loader module ``hamt_b200.py`` at the repo root.
"""
from .config import HamtConfig, ALL_TASKS  # noqa: F401

__all__ = ["HamtConfig", "ALL_TASKS"]
