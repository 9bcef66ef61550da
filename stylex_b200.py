"""Import alias: ``import stylex_b200`` -> the package directory
``explaining-in-style-reproducibility-study_b200/`` (whose name is not a Python identifier)."""
import importlib
import sys

_pkg = importlib.import_module("explaining-in-style-reproducibility-study_b200")
sys.modules[__name__] = _pkg
