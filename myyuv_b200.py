"""Import alias: the package directory `yuv-manipulations-2_b200` is not a valid Python identifier."""
import importlib as _il
import sys as _sys

_pkg = _il.import_module("yuv-manipulations-2_b200")
_sys.modules[__name__] = _pkg
