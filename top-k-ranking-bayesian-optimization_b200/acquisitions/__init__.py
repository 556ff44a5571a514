"""Acquisition functions on the hot path (same sub-module names as the reference's acquisitions/__init__.py:1;
``ei`` and ``indiff_pes`` are outside this package's scope)."""
from . import pes, dts, rank_pes  # noqa: F401
