"""Importable alias of the package directory ``audio-deepfake-explainability_b200/`` (a hyphen is not a
valid Python identifier, so ``import audio_deepfake_explainability_b200`` resolves here and this shim
redirects the package search path to the hyphenated source directory)."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "audio-deepfake-explainability_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py"), encoding="utf-8") as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
