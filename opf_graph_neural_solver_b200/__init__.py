"""Importable alias of the hyphenated source directory ``opf-graph-neural-solver_b200/``.

``import opf_graph_neural_solver_b200 as gns`` loads the real package that lives next to
this directory (its name carries a hyphen and cannot be imported directly).
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "opf-graph-neural-solver_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f, _real
