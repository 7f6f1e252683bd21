"""Importable alias of the package directory ``puc-fluidsimulation-project_b200/``
(hyphens cannot appear in a Python import name).  All code lives there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "puc-fluidsimulation-project_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
