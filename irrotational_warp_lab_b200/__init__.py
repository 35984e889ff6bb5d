"""Import shim: ``irrotational-warp-lab_b200/`` (the package directory the build contract
names) is not a valid Python identifier, so this package forwards to it."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "irrotational-warp-lab_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
