"""Import shim: the product package lives in the directory ``visual-odom-pipeline_b200/`` (a name
Python cannot import directly); this makes it importable as ``visual_odom_pipeline_b200``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "visual-odom-pipeline_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
