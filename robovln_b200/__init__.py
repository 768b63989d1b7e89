"""Import alias for the package that lives in ``robo-vln_b200/`` (a hyphen cannot appear in a
Python module name).  ``import robovln_b200`` executes that package's ``__init__`` with this
module's ``__path__`` pointing at the real directory, so sub-modules resolve there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "robo-vln_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
