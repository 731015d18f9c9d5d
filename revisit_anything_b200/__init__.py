"""Importable alias for the package directory `revisit-anything_b200/` (a hyphen is not a valid
Python identifier, the directory name is fixed by the build contract).  `import revisit_anything_b200`
and every `revisit_anything_b200.<submodule>` resolve to files inside `revisit-anything_b200/`."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "revisit-anything_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
