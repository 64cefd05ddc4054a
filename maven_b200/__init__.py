"""Importable alias for the package directory `multimodal-supernovae_b200/` (a hyphen is not a
valid Python identifier).  `import maven_b200` executes that directory's __init__ and resolves
submodules (maven_b200.transformer_utils, maven_b200.loss, ...) from it."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "multimodal-supernovae_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
