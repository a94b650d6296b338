"""Import alias for the product package.

The product lives in the directory ``hypatia.jl_b200/`` (the name the build contract asks
for).  A directory name containing a dot cannot be imported with a plain ``import``
statement, so this two-line package points its ``__path__`` at that directory:
``import hypatia_b200.capi`` loads ``hypatia.jl_b200/capi.py``.
"""
import os as _os

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
__path__ = [_os.path.join(_ROOT, "hypatia.jl_b200")]
PACKAGE_DIR = __path__[0]
