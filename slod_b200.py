"""Import alias for the hyphenated package directory ``dealii-slod_b200``."""
import importlib as _il
import sys as _sys

_pkg = _il.import_module("dealii-slod_b200")
_sys.modules[__name__] = _pkg
