"""Importable alias of the product package.

The package directory is named ``synthetic-audio-detection_b200`` (not a valid Python identifier), so this
module re-exports it as ``sad_b200``: ``import sad_b200.inference_runner`` loads
``synthetic-audio-detection_b200/inference_runner.py``.
"""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "synthetic-audio-detection_b200")
__path__ = [_REAL]
with open(_os.path.join(_REAL, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_REAL, "__init__.py"), "exec"))
