"""Import shim: the package directory is named ``kb-ner_b200`` (not a valid Python identifier),
so ``import kbner_b200`` resolves here and executes the real package in place."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "kb-ner_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
