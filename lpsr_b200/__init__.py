"""Import shim: ``import lpsr_b200`` -> the package in ``license-plate-detection-and-recognition-with-image-enhancement_b200/``
(a directory name with hyphens cannot be imported directly)."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "license-plate-detection-and-recognition-with-image-enhancement_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
