"""Import shim: the product package lives in the directory ``gan-inpainting_b200/`` (the
reference repo's name, which is not a valid Python identifier).  ``import gan_inpainting_b200``
resolves every submodule from there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "gan-inpainting_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
