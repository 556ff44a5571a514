"""Acquisition functions on the hot path (same sub-module names as the reference's acquisitions/__init__.py:1;
``ei`` is outside this package's scope)."""
from . import pes, dts, rank_pes, indiff_pes  # noqa: F401
