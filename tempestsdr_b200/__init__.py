"""Import shim: the package directory is named after the reference
(`tempestsdr.jl_b200/`, not a valid Python identifier), so this module points
its search path there and re-exports the host API.

    import tempestsdr_b200 as tsdr
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "tempestsdr.jl_b200")
__path__.insert(0, _PKG_DIR)

from .api import *  # noqa: E402,F401,F403
from .api import __all__  # noqa: E402,F401
