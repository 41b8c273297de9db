"""Importable alias of the package directory
``seamless-through-breaking-rethinking-image-stitching-for-optimal-alignment_b200/``
(its name is fixed by the project layout and is not a valid Python identifier).
``import stitch_b200`` executes that directory's ``__init__`` as this package."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "seamless-through-breaking-rethinking-image-stitching-for-optimal-alignment_b200")
__path__[:] = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"), globals())
