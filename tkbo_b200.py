"""Importable alias: ``import tkbo_b200`` loads the package directory
``top-k-ranking-bayesian-optimization_b200`` (a hyphenated name, which the reference's scripts load with
importlib.import_module exactly like their own package; experiments/MPES/MPES_Forr.py:22-24)."""
import importlib
import sys

_pkg = importlib.import_module("top-k-ranking-bayesian-optimization_b200")
sys.modules[__name__] = _pkg
