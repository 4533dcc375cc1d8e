"""Importable alias of the package directory `halo2-plonky2-verifier_b200/` (its name is not a valid Python
identifier). `import b200zk` executes that package's __init__ under this name."""
import os as _os

_pkg = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "halo2-plonky2-verifier_b200")
__path__ = [_pkg]
with open(_os.path.join(_pkg, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_pkg, "__init__.py"), "exec"))
