"""Import alias: the package directory is `pytorch-pose-estimation_b200/` (not a valid Python identifier),
so this shim makes it importable as `pose_b200` by pointing the package search path at it."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "pytorch-pose-estimation_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
